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

// ---- ConvTemporalGraphical as ONE temporal convolution --------------------------------------------------------------
// The (Kt x 1) convolution Cin -> K*C and the adjacency contraction are both linear, so their composition is a Conv1d
// over time on the channels-last rows x[n, t, v*Cin + ci] -> y[n, t, w*C + c]:
//   Weff[w*C + c][v*Cin + ci][dt] = sum_k A[k][v][w] * W[k*C + c][ci][dt]
//   beff[w*C + c]                 = sum_k (sum_v A[k][v][w]) * b[k*C + c]
// (zero padding in time commutes with the contraction: the bias is added at every output position either way).  The
// [N, T, V, K*C] intermediate (25 MB at 256 clips for the 9-joint graph) is never formed; the composition is redone every
// step because W changes every step (35 k / 62 k elements).
__global__ void __launch_bounds__(256) gcn_compose_fwd_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                              const float* __restrict__ A, float* __restrict__ Weff,
                                                              float* __restrict__ beff, int V, int K, int C, int Cin,
                                                              int Kt) {
  __shared__ float sA[kMaxA];
  for (int i = threadIdx.x; i < K * V * V; i += blockDim.x) sA[i] = A[i];
  __syncthreads();
  const int VC = V * C, VCin = V * Cin;
  const long total = (long)VC * VCin * Kt;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int dt = (int)(i % Kt); const int col = (int)((i / Kt) % VCin); const int row = (int)(i / ((long)Kt * VCin));
    const int w = row / C, c = row - w * C, v = col / Cin, ci = col - v * Cin;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(sA[(k * V + v) * V + w], __ldg(W + ((long)(k * C + c) * Cin + ci) * Kt + dt), acc);
    Weff[i] = acc;
  }
  if (blockIdx.x == 0 && beff != nullptr)
    for (int row = threadIdx.x; row < VC; row += blockDim.x) {
      const int w = row / C, c = row - w * C;
      float acc = 0.f;
      if (b != nullptr)
        for (int k = 0; k < K; ++k) {
          float a = 0.f;
          for (int v = 0; v < V; ++v) a += sA[(k * V + v) * V + w];
          acc = fmaf(a, b[k * C + c], acc);
        }
      beff[row] = acc;
    }
}

// dW[k*C + c][ci][dt] += sum_{v,w} A[k][v][w] * dWeff[w*C + c][v*Cin + ci][dt];  db[k*C + c] += sum_w (sum_v A[k][v][w]) dbeff[w*C + c]
__global__ void __launch_bounds__(256) gcn_compose_bwd_kernel(const float* __restrict__ dWeff, const float* __restrict__ dbeff,
                                                              const float* __restrict__ A, float* __restrict__ dW,
                                                              float* __restrict__ db, int V, int K, int C, int Cin, int Kt,
                                                              int layout) {
  __shared__ float sA[kMaxA];
  for (int i = threadIdx.x; i < K * V * V; i += blockDim.x) sA[i] = A[i];
  __syncthreads();
  const int VCin = V * Cin;
  const long total = (long)K * C * Cin * Kt;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int dt = (int)(i % Kt); const int ci = (int)((i / Kt) % Cin); const int kc = (int)(i / ((long)Kt * Cin));
    const int k = kc / C, c = kc - k * C;
    float acc = 0.f;
    for (int v = 0; v < V; ++v)
      for (int w = 0; w < V; ++w)
        // layout 0: dWeff[w*C + c][v*Cin + ci][dt] (the convolution's own weight layout);
        // layout 1: dWeffT[dt*VCin + v*Cin + ci][w*C + c] (s2ag_window_wgrad: window column major)
        acc = fmaf(sA[(k * V + v) * V + w],
                   __ldg(layout == 0 ? dWeff + ((long)(w * C + c) * VCin + v * Cin + ci) * Kt + dt
                                     : dWeff + ((long)dt * VCin + v * Cin + ci) * (V * C) + (w * C + c)), acc);
    dW[i] += acc;
  }
  if (blockIdx.x == 0 && db != nullptr && dbeff != nullptr)
    for (int kc = threadIdx.x; kc < K * C; kc += blockDim.x) {
      const int k = kc / C, c = kc - k * C;
      float acc = 0.f;
      for (int w = 0; w < V; ++w) {
        float a = 0.f;
        for (int v = 0; v < V; ++v) a += sA[(k * V + v) * V + w];
        acc = fmaf(a, dbeff[w * C + c], acc);
      }
      db[kc] += acc;
    }
}
}  // namespace

extern "C" int s2ag_gcn_compose_fwd(const float* W, const float* b, const float* A, float* Weff, float* beff, int V, int K,
                                    int C, int Cin, int Kt, void* stream) {
  S2AG_CHECK_ARG(W && A && Weff && V > 0 && K > 0 && C > 0 && Cin > 0 && Kt > 0 && K * V * V <= kMaxA);
  long total = (long)V * C * V * Cin * Kt;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 4) blocks = 148 * 4;
  auto kfn = &gcn_compose_fwd_kernel;
  S2AG_LAUNCH(kfn, blocks, 256, 0, stream, W, b, A, Weff, beff, V, K, C, Cin, Kt);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
extern "C" int s2ag_gcn_compose_bwd(const float* dWeff, const float* dbeff, const float* A, float* dW, float* db, int V,
                                    int K, int C, int Cin, int Kt, int layout, void* stream) {
  S2AG_CHECK_ARG(dWeff && A && dW && V > 0 && K > 0 && C > 0 && Cin > 0 && Kt > 0 && K * V * V <= kMaxA);
  long total = (long)K * C * Cin * Kt;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 4) blocks = 148 * 4;
  auto kfn = &gcn_compose_bwd_kernel;
  S2AG_CHECK_ARG(layout == 0 || layout == 1);
  S2AG_LAUNCH(kfn, blocks, 256, 0, stream, dWeff, dbeff, A, dW, db, V, K, C, Cin, Kt, layout);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

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
