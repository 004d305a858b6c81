// Engine selection for the dense contractions of the hot path: the tcgen05 kernel (gemm_umma.cuh)
// wherever a 128 x BN tensor-core tile is worthwhile, the exact-fp32 SIMT kernel (gemm_simt.cuh)
// for the tiny shapes (K < 32, N < 16, M < 64) and in the CPU logic-emulation build of tests/emu.
// s2ag_set_engine(1) forces the SIMT kernel everywhere (A/B parity checks on the GPU).
#pragma once
#include "gemm_simt.cuh"
#include "gemm_umma.cuh"

namespace s2ag {
extern int g_engine;  // 0 = auto, 1 = SIMT only

template <class LdA, class LdB, class Epi>
static inline void launch_gemm(const LdA& a, const LdB& b, const Epi& epi, int M, int N, int K, int nbatch, int splitk,
                               void* stream) {
  if (M <= 0 || N <= 0) return;
  if (splitk < 1) splitk = 1;
#ifndef S2AG_EMU
  if (g_engine == 0 && umma::worthwhile(M, N, K)) {
    umma::launch(a, b, epi, M, N, K, nbatch, splitk, stream);
    return;
  }
#endif
  launch_gemm_simt(a, b, epi, M, N, K, nbatch, splitk, stream);
}
}  // namespace s2ag
