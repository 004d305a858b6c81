// ST-GCN adjacency contraction, channels-last.
// Replaces torch.einsum('nkctv,kvw->nctw') in ConvTemporalGraphical.forward (reference
// net/utils/tgcn.py:66-69).  x[M,V,K*C] -> y[M,V,C] with A[K,V,V] held in shared memory.
// Tiny contraction dims (K*V <= 45): SIMT, one thread per output element, coalesced over c.
#include "s2ag.h"
#include "common.cuh"

namespace {
constexpr int kMaxA = 5 * 9 * 9;

__global__ void __launch_bounds__(256) graph_fwd_kernel(const float* __restrict__ x, const float* __restrict__ A,
                                                        float* __restrict__ y, int M, int V, int K, int C) {
  __shared__ float sA[kMaxA];
  for (int i = threadIdx.x; i < K * V * V; i += blockDim.x) sA[i] = A[i];
  __syncthreads();
  const int VC = V * C, KC = K * C;
  const long total = (long)M * VC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); const int w = (int)((i / C) % V); const long m = i / VC;
    const float* xr = x + m * (long)V * KC;
    float acc = 0.f;
    for (int k = 0; k < K; ++k)
      for (int v = 0; v < V; ++v) acc = fmaf(__ldg(xr + v * KC + k * C + c), sA[(k * V + v) * V + w], acc);
    y[i] = acc;
  }
}

__global__ void __launch_bounds__(256) graph_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ A,
                                                        float* __restrict__ dx, int M, int V, int K, int C) {
  __shared__ float sA[kMaxA];
  for (int i = threadIdx.x; i < K * V * V; i += blockDim.x) sA[i] = A[i];
  __syncthreads();
  const int VC = V * C, KC = K * C;
  const long total = (long)M * V * KC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int kc = (int)(i % KC); const int v = (int)((i / KC) % V); const long m = i / ((long)V * KC);
    const int k = kc / C, c = kc % C;
    const float* dr = dy + m * VC;
    float acc = 0.f;
    for (int w = 0; w < V; ++w) acc = fmaf(__ldg(dr + w * C + c), sA[(k * V + v) * V + w], acc);
    dx[i] = acc;
  }
}
}  // namespace

extern "C" int s2ag_graph_fwd(const float* x, const float* A, float* y, int M, int V, int K, int C, void* stream) {
  S2AG_CHECK_ARG(x && A && y && M >= 0 && V > 0 && K > 0 && C > 0 && K * V * V <= kMaxA);
  long total = (long)M * V * C;
  if (total == 0) return S2AG_OK;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  auto kfn = &graph_fwd_kernel;
  S2AG_LAUNCH(kfn, blocks, 256, 0, stream, x, A, y, M, V, K, C);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
extern "C" int s2ag_graph_bwd(const float* dy, const float* A, float* dx, int M, int V, int K, int C, void* stream) {
  S2AG_CHECK_ARG(dy && A && dx && M >= 0 && V > 0 && K > 0 && C > 0 && K * V * V <= kMaxA);
  long total = (long)M * V * K * C;
  if (total == 0) return S2AG_OK;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  auto kfn = &graph_bwd_kernel;
  S2AG_LAUNCH(kfn, blocks, 256, 0, stream, dy, A, dx, M, V, K, C);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
