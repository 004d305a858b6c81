// Engine selection for the dense contractions of the hot path: the tcgen05 kernel (gemm_umma.cuh)
// wherever a 128 x BN tensor-core tile is worthwhile, the exact-fp32 SIMT kernel (gemm_simt.cuh)
// for the tiny shapes (K < 32, N < 16, M < 64) and in the CPU logic-emulation build of tests/emu.
// s2ag_set_engine(1) forces the SIMT kernel everywhere (A/B parity checks on the GPU).
#pragma once
#include <cstdio>
#include "gemm_simt.cuh"
#include "gemm_umma.cuh"
#include "gemm_umma_packed.cuh"
#include "gemm_umma_tt.cuh"

namespace s2ag {
extern int g_engine;  // 0 = auto, 1 = SIMT only
void* scratch_get(void* stream, long bytes);  // capi.cu: scratch registered for `stream` if it holds >= bytes, else NULL

template <class LdA, class LdB, class Epi>
static inline void launch_gemm_untimed(const LdA& a, const LdB& b, const Epi& epi, int M, int N, int K, int nbatch,
                                       int splitk, void* stream) {
#ifndef S2AG_EMU
  if (g_engine == 0 && umma::worthwhile(M, N, K)) {
    // B reused by >= 4 row tiles (a weight, or any operand small next to A): pack it once into the tensor-core
    // operand image and let the CTAs fetch it by TMA (gemm_umma_packed.cuh) -- needs a scratch buffer registered
    // for this stream (s2ag_register_scratch); without one the on-the-fly kernel runs.
    // big weight-side contractions: both operands pre-packed, 256 x BN tiles, no conversion in the main loop
    // (gemm_umma_tt.cuh); needs scratch for both operand images on this stream
    if (!(umma::g_dbg_flags & 256) && umma::launch_tt(a, b, epi, M, N, K, nbatch, splitk, &scratch_get, stream)) return;
    if (M >= 4 * umma::BM && !(umma::g_dbg_flags & 256)) {
      const long bytes = umma::packed_bytes(N, K, nbatch);
      void* img = scratch_get(stream, bytes);
      if (img != nullptr) {
        const umma::PackedB pb = umma::pack_operand(b, N, K, nbatch, img, stream);
        umma::launch_packed(a, pb, epi, M, N, K, nbatch, splitk, stream);
        return;
      }
    }
    umma::launch(a, b, epi, M, N, K, nbatch, splitk, stream);
    return;
  }
#endif
  launch_gemm_simt(a, b, epi, M, N, K, nbatch, splitk, stream);
}

template <class LdA, class LdB, class Epi>
static inline void launch_gemm(const LdA& a, const LdB& b, const Epi& epi, int M, int N, int K, int nbatch, int splitk,
                               void* stream) {
  if (M <= 0 || N <= 0) return;
  if (splitk < 1) splitk = 1;
#ifndef S2AG_EMU
  if (umma::g_dbg_flags & 32) {
    // bring-up aid (s2ag_debug_flags bit 5, eager mode only): one stderr line per contraction with its device time
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, (cudaStream_t)stream);
    launch_gemm_untimed(a, b, epi, M, N, K, nbatch, splitk, stream);
    cudaEventRecord(e1, (cudaStream_t)stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[gemm] %8.1f us  M=%d N=%d K=%d batch=%d splitk=%d  %6.1f TF/s  %s\n", ms * 1e3f, M, N, K, nbatch, splitk,
            2.0 * M * N * (double)K * nbatch / (ms * 1e-3) * 1e-12, __PRETTY_FUNCTION__);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return;
  }
#endif
  launch_gemm_untimed(a, b, epi, M, N, K, nbatch, splitk, stream);
}

}  // namespace s2ag
