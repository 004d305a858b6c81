// Micro-probe: aggregate L2 -> SM read bandwidth on this part for the access patterns of the hot path.
//   mode 0: every CTA streams a DISJOINT slice of an L2-resident buffer (LDG.256)
//   mode 1: every CTA streams the SAME L2-resident buffer (all SMs read the same lines: operand-tile / h-image pattern)
//   mode 2: like 1 with TMA bulk copies (cp.async.bulk global -> shared), 16 KB per copy
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(512) read_kernel(const float* __restrict__ buf, size_t floats_per_cta, int shared_mode, int iters, float* out) {
  const float* base = shared_mode ? buf : buf + (size_t)blockIdx.x * floats_per_cta;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    for (size_t i = (size_t)threadIdx.x * 8; i + 8 <= floats_per_cta; i += 512 * 8) {
      float v[8];
      asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]),"=f"(v[1]),"=f"(v[2]),"=f"(v[3]),"=f"(v[4]),"=f"(v[5]),"=f"(v[6]),"=f"(v[7]) : "l"(base + i));
      acc += v[0] + v[7];
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(128) tma_kernel(const unsigned char* __restrict__ buf, size_t bytes_per_cta, int shared_mode, int iters, float* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t bar = sbase;                 // 4 stage barriers at sbase + 8*s
  const uint32_t data = sbase + 128;
  const int STAGES = 4; const uint32_t CH = 16384;
  const unsigned char* base = shared_mode ? buf : buf + (size_t)blockIdx.x * bytes_per_cta;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar + 8 * s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t nch = bytes_per_cta / CH;
    size_t issued = 0, done = 0; const size_t total = nch * iters;
    uint32_t phase[4] = {0, 0, 0, 0};
    while (done < total) {
      while (issued < total && issued < done + STAGES) {
        const int s = issued % STAGES;
        const unsigned char* src = base + (issued % nch) * CH;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar + 8 * s), "r"(CH) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(data + s * CH), "l"(src), "r"(CH), "r"(bar + 8 * s) : "memory");
        ++issued;
      }
      const int s = done % STAGES;
      uint32_t ok = 0;
      while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar + 8 * s), "r"(phase[s]) : "memory");
      phase[s] ^= 1; ++done;
    }
  }
  __syncthreads();
  if (sm[200] == 77 && out) out[1] = 1.f;
}

int main() {
  const size_t total_bytes = 64u << 20;  // 64 MB: L2 resident (126 MB L2)
  float* buf; float* out;
  CK(cudaMalloc(&buf, total_bytes)); CK(cudaMalloc(&out, 64)); CK(cudaMemset(buf, 0, total_bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas : {76, 148}) {
    for (int mode = 0; mode < 2; ++mode) {
      // per-CTA slice: 256 KB (h-image-like) ; shared mode: all CTAs read the same 256 KB
      const size_t fl = (256u << 10) / 4; const int iters = 64;
      read_kernel<<<ctas, 512>>>(buf, fl, mode, 2, out);  // warm (brings the lines into L2)
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); read_kernel<<<ctas, 512>>>(buf, fl, mode, iters, out); cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double gb = (double)ctas * fl * 4 * iters / 1e9;
      printf("LDG.256  ctas=%3d %s 256KB/CTA : %.2f TB/s aggregate, %.1f B/clk/SM @1.9GHz\n", ctas, mode ? "SAME lines " : "disjoint   ", gb / ms, gb / ms * 1e12 / 1e3 / ctas / 1.9e9 * 1e0);
    }
    for (int mode = 0; mode < 2; ++mode) {
      const size_t by = 256u << 10; const int iters = 64;
      cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 4 * 16384);
      tma_kernel<<<ctas, 128, 128 + 4 * 16384>>>((const unsigned char*)buf, by, mode, 2, out);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); tma_kernel<<<ctas, 128, 128 + 4 * 16384>>>((const unsigned char*)buf, by, mode, iters, out); cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double gb = (double)ctas * by * iters / 1e9;
      printf("TMA bulk ctas=%3d %s 256KB/CTA : %.2f TB/s aggregate, %.1f B/clk/SM @1.9GHz\n", ctas, mode ? "SAME lines " : "disjoint   ", gb / ms, gb / ms * 1e12 / 1e3 / ctas / 1.9e9);
    }
  }
  return 0;
}
