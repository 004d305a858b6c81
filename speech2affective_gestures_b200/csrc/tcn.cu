// Text-encoder path: nn.Embedding(+Dropout), old-style weight_norm, and the causal dilated
// TemporalBlock (kernel 2) as a conv-as-GEMM with K = 2*C over channels-last [B,T,C].
// Reference: net/tcn.py:7-46 (Chomp1d/TemporalBlock), net/multimodal_context_net_v2.py:61-91.
// The chomped (acausal) columns the reference computes and discards (tcn.py:12-13) are never
// computed here, and the two taps are gathered by the operand loader, so no padded copy exists.
#include "s2ag.h"
#include "gemm.cuh"

using namespace s2ag;
namespace s2ag { void launch_colsum(const float* dy, long ld, float* db, int M, int N, void* stream); }

namespace {

// ---------------------------------------------------------------- weight_norm
__global__ void __launch_bounds__(128) weight_norm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                              float* __restrict__ w, float* __restrict__ norm, int Ci,
                                                              int k) {
  __shared__ float red[4];
  const int co = blockIdx.x, n = Ci * k;
  const float* vr = v + (long)co * n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 128) { float t = vr[i]; s = fmaf(t, t, s); }
  s = s2ag_warp_sum(s);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
  __syncthreads();
  const float nrm = sqrtf(red[0] + red[1] + red[2] + red[3]);
  const float sc = g[co] / nrm;
  for (int i = threadIdx.x; i < n; i += 128) {
    const int c = i / k, j = i % k;  // v is [Ci][k]; w is [k][Ci]
    w[(long)co * n + j * Ci + c] = vr[i] * sc;
  }
  if (threadIdx.x == 0) norm[co] = nrm;
}

__global__ void __launch_bounds__(128) weight_norm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                              const float* __restrict__ g,
                                                              const float* __restrict__ norm, float* __restrict__ dv,
                                                              float* __restrict__ dg, int Ci, int k) {
  __shared__ float red[4];
  const int co = blockIdx.x, n = Ci * k;
  const float* vr = v + (long)co * n;
  const float* dwr = dw + (long)co * n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 128) {
    const int c = i / k, j = i % k;
    s = fmaf(dwr[j * Ci + c], vr[i], s);
  }
  s = s2ag_warp_sum(s);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
  __syncthreads();
  const float dot = red[0] + red[1] + red[2] + red[3];
  const float nrm = norm[co], gg = g[co];
  const float a = gg / nrm, b = gg * dot / (nrm * nrm * nrm);
  for (int i = threadIdx.x; i < n; i += 128) {
    const int c = i / k, j = i % k;
    dv[(long)co * n + i] += a * dwr[j * Ci + c] - b * vr[i];
  }
  if (threadIdx.x == 0) dg[co] += dot / nrm;
}

// ---------------------------------------------------------------- TCN epilogues
// y = drop(relu(acc + bias)); optionally out = relu(y + res)
struct EpiTcn {
  float* y; float* out; const float* bias; const float* res; long ld; float p; unsigned long long seed;
  const unsigned long long* seed_dev;
  __device__ __forceinline__ void operator()(int, int m, int n, float acc, bool) const {
    const long i = (long)m * ld + n;
    float v = acc + __ldg(bias + n);
    v = v > 0.f ? v : 0.f;
    if (p > 0.f) v *= s2ag_dropout_scale(seed + (seed_dev ? seed_dev[0] : 0ull), (unsigned long long)i, p);
    y[i] = v;
    if (out) { float o = v + __ldg(res + i); out[i] = o > 0.f ? o : 0.f; }
  }
  struct Col { int n; float bias; unsigned long long seed; };
  __device__ __forceinline__ Col col(int, int n) const {
    return Col{n, __ldg(bias + n), seed + ((p > 0.f && seed_dev) ? seed_dev[0] : 0ull)};
  }
  __device__ __forceinline__ void apply(const Col& c, int m, float acc, bool) const {
    const long i = (long)m * ld + c.n;
    float v = acc + c.bias;
    v = v > 0.f ? v : 0.f;
    if (p > 0.f) v *= s2ag_dropout_scale(c.seed, (unsigned long long)i, p);
    y[i] = v;
    if (out) { float o = v + __ldg(res + i); out[i] = o > 0.f ? o : 0.f; }
  }
};

// g_out = dout * (out>0); dx = g_out; g2 = g_out * (y2>0) * scale
__global__ void tcn_bwd_head_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                    const float* __restrict__ y2, float* __restrict__ dx, float* __restrict__ g2,
                                    long n, float scale) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float go = out[i] > 0.f ? dout[i] : 0.f;
    dx[i] = go;
    g2[i] = y2[i] > 0.f ? go * scale : 0.f;
  }
}

// ---------------------------------------------------------------- embedding
__global__ void embedding_fwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ table,
                                     float* __restrict__ out, long ldo, long n, int D, long V, float p,
                                     unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev) seed += seed_dev[0];
  const long total = n * D;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long i = e / D; const int d = (int)(e % D);
    const long r = idx[i];
    if (r < 0 || r >= V) S2AG_DEVICE_TRAP();  // nn.Embedding device-asserts here too
    out[i * ldo + d] = __ldg(table + r * (long)D + d) * s2ag_dropout_scale(seed, (unsigned long long)e, p);
  }
}
// dtable[idx[i]] += dout[i] * dropout mask.  Most positions of a clip hold the SAME token (24 of the 34 frames are the
// padding id: extend_word_seq leaves zeros between the ~10 words), so per-element atomics serialise thousands of adds on
// the 300 addresses of one table row (72 us at 256 clips).  A block owns a chunk of consecutive rows and a thread one
// feature column: it walks the chunk keeping a running sum while the index repeats... and, because the repeats are
// interleaved with the words, also a second running sum for the chunk's HOT index (taken from the chunk's first rows):
// rows of the hot index never touch memory until the chunk is done.  Any index may be hot; only the speed depends on it.
constexpr int kEmbRows = 64;
__global__ void __launch_bounds__(128) embedding_bwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dout,
                                                            long ldo, float* __restrict__ dtable, long n, int D, long V,
                                                            float p, unsigned long long seed,
                                                            const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev) seed += seed_dev[0];
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  if (d >= D) return;   // no barriers in this kernel
  for (long i0 = (long)blockIdx.x * kEmbRows; i0 < n; i0 += (long)gridDim.x * kEmbRows) {
    const long i1 = i0 + kEmbRows < n ? i0 + kEmbRows : n;
    long hot = idx[i0];   // the more frequent of the chunk's first rows
    if (i0 + 2 < i1 && idx[i0 + 1] == idx[i0 + 2]) hot = idx[i0 + 1];
    float hot_sum = 0.f;
    for (long i = i0; i < i1; ++i) {
      const long r = idx[i];
      if (r < 0 || r >= V) S2AG_DEVICE_TRAP();
      const float gsc = s2ag_dropout_scale(seed, (unsigned long long)(i * D + d), p);
      if (gsc == 0.f) continue;
      const float v = dout[i * ldo + d] * gsc;
      if (r == hot) hot_sum += v;
      else atomicAdd(dtable + r * (long)D + d, v);
    }
    if (hot_sum != 0.f) atomicAdd(dtable + hot * (long)D + d, hot_sum);
  }
}

static inline LdConv<ORDER_KKC> tcn_loader(const float* x, int T, int C, int d, int sgn) {
  // rows (b,t); K = (tap j, channel c); forward: source t + (j-1)*d ; data-grad: source t + (1-j)*d
  return LdConv<ORDER_KKC>{x, T, 1, C, T, 1, 2, 1, 1, 1, d, 1, sgn, sgn > 0 ? -d : d, 0, (long)C};
}

}  // namespace

extern "C" int s2ag_weight_norm_fwd(const float* v, const float* g, float* w, float* norm, int Co, int Ci, int k,
                                    void* stream) {
  S2AG_CHECK_ARG(v && g && w && norm && Co > 0 && Ci > 0 && k > 0);
  auto kfn = &weight_norm_fwd_kernel;
  S2AG_LAUNCH(kfn, Co, 128, 0, stream, v, g, w, norm, Ci, k);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
extern "C" int s2ag_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* norm,
                                    float* dv, float* dg, int Co, int Ci, int k, void* stream) {
  S2AG_CHECK_ARG(dw && v && g && norm && dv && dg && Co > 0 && Ci > 0 && k > 0);
  auto kfn = &weight_norm_bwd_kernel;
  S2AG_LAUNCH(kfn, Co, 128, 0, stream, dw, v, g, norm, dv, dg, Ci, k);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_tcn_block_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                  float* y1, float* y2, float* out, int B, int T, int C, int dilation,
                                  float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  S2AG_CHECK_ARG(x && w1 && b1 && w2 && b2 && y1 && y2 && out && B >= 0 && T > 0 && C > 0 && dilation > 0);
  S2AG_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f);
  const int M = B * T, K = 2 * C;
  LdPlain<true> bw1{w1, (long)K, 1, 0}, bw2{w2, (long)K, 1, 0};
  EpiTcn e1{y1, nullptr, b1, nullptr, (long)C, p_drop, seed, (const unsigned long long*)seed_dev};
  launch_gemm(tcn_loader(x, T, C, dilation, +1), bw1, e1, M, C, K, 1, 1, stream);
  EpiTcn e2{y2, out, b2, x, (long)C, p_drop, seed + 0x1234567ull, (const unsigned long long*)seed_dev};
  launch_gemm(tcn_loader(y1, T, C, dilation, +1), bw2, e2, M, C, K, 1, 1, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

// phases: 1 = data path (head kernel, g1 = dgrad(g2; W2), dx += dgrad(g1; W1)), 2 = parameter path (dw2, db2, dw1, db1;
// reads the g2 / g1 planes phase 1 left in ws), 3 = both.  A caller may run phase 2 on another stream, ordered after
// phase 1 by an event, so that the weight gradients leave the data-gradient chain of the text encoder.
extern "C" int s2ag_tcn_block_bwd_phased(const float* dout, const float* x, const float* y1, const float* y2,
                                         const float* out, const float* w1, const float* w2, float* dx, float* dw1,
                                         float* db1, float* dw2, float* db2, float* ws, int B, int T, int C,
                                         int dilation, float p_drop, int phases, void* stream) {
  S2AG_CHECK_ARG(x && y1 && w1 && w2 && ws && (phases & 3) && !(phases & ~3));
  S2AG_CHECK_ARG(!(phases & 1) || (dout && y2 && out && dx));
  S2AG_CHECK_ARG(!(phases & 2) || (dw1 && db1 && dw2 && db2));
  S2AG_CHECK_ARG(B >= 0 && T > 0 && C > 0 && dilation > 0 && p_drop >= 0.f && p_drop < 1.f);
  const int M = B * T, K = 2 * C;
  if (M == 0) return S2AG_OK;
  const long n = (long)M * C;
  const float scale = 1.f / (1.f - p_drop);
  float* g2 = ws; float* g1 = ws + n;
  if (phases & 1) {
    {
      int blocks = (int)((n + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
      auto kfn = &tcn_bwd_head_kernel;
      S2AG_LAUNCH(kfn, blocks, 256, 0, stream, dout, out, y2, dx, g2, n, scale);
    }
    // g1 = dgrad(g2; W2) * relu'(y1) * scale
    {
      LdWdgrad<ORDER_KKC> b{w2, C, 2, (long)K, 1, (long)C};
      EpiGeneric e = make_epi(g1, (long)C);
      e.alpha = scale; e.mul_src = y1; e.ld_mul = C; e.mul_act = S2AG_ACT_RELU;
      launch_gemm(tcn_loader(g2, T, C, dilation, -1), b, e, M, C, K, 1, 1, stream);
    }
    // dx += dgrad(g1; W1)
    {
      LdWdgrad<ORDER_KKC> b{w1, C, 2, (long)K, 1, (long)C};
      launch_gemm(tcn_loader(g1, T, C, dilation, -1), b, make_epi(dx, (long)C, nullptr, 0, 0.f, 1), M, C, K, 1, 1,
                  stream);
    }
  }
  if (phases & 2) {
    const int sk = pick_splitk(C, K, M, 1);
    // conv2 weight / bias gradients
    {
      LdPlain<false> a{g2, 1, (long)C, 0};
      LdT<LdConv<ORDER_KKC>> b{tcn_loader(y1, T, C, dilation, +1)};
      launch_gemm(a, b, make_epi(dw2, (long)K, nullptr, 0, 0.f, sk > 1 ? 2 : 1), C, K, M, 1, sk, stream);
      launch_colsum(g2, C, db2, M, C, stream);
    }
    // conv1 weight / bias gradients
    {
      LdPlain<false> a{g1, 1, (long)C, 0};
      LdT<LdConv<ORDER_KKC>> b{tcn_loader(x, T, C, dilation, +1)};
      launch_gemm(a, b, make_epi(dw1, (long)K, nullptr, 0, 0.f, sk > 1 ? 2 : 1), C, K, M, 1, sk, stream);
      launch_colsum(g1, C, db1, M, C, stream);
    }
  }
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_tcn_block_bwd(const float* dout, const float* x, const float* y1, const float* y2,
                                  const float* out, const float* w1, const float* w2, float* dx, float* dw1,
                                  float* db1, float* dw2, float* db2, float* ws, int B, int T, int C, int dilation,
                                  float p_drop, void* stream) {
  S2AG_CHECK_ARG(dout && x && y1 && y2 && out && w1 && w2 && dx && dw1 && db1 && dw2 && db2 && ws);
  return s2ag_tcn_block_bwd_phased(dout, x, y1, y2, out, w1, w2, dx, dw1, db1, dw2, db2, ws, B, T, C, dilation, p_drop,
                                   3, stream);
}

extern "C" int s2ag_embedding_fwd(const int64_t* idx, const float* table, float* out, long ldo, long n, int D, long V,
                                  float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  S2AG_CHECK_ARG(idx && table && out && n >= 0 && D > 0 && V > 0 && ldo >= D && p_drop >= 0.f && p_drop < 1.f);
  long total = n * D;
  if (total == 0) return S2AG_OK;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  auto kfn = &embedding_fwd_kernel;
  S2AG_LAUNCH(kfn, blocks, 256, 0, stream, idx, table, out, ldo, n, D, V, p_drop, (unsigned long long)seed,
              (const unsigned long long*)seed_dev);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
extern "C" int s2ag_embedding_bwd(const int64_t* idx, const float* dout, long ldo, float* dtable, long n, int D, long V,
                                  float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  S2AG_CHECK_ARG(idx && dout && dtable && n >= 0 && D > 0 && V > 0 && ldo >= D && p_drop >= 0.f && p_drop < 1.f);
  long total = n * D;
  if (total == 0) return S2AG_OK;
  long chunks = (n + kEmbRows - 1) / kEmbRows; if (chunks > 148 * 16) chunks = 148 * 16;
  auto kfn = &embedding_bwd_kernel;
  S2AG_LAUNCH(kfn, dim3((unsigned)chunks, (unsigned)((D + 127) / 128)), 128, 0, stream, idx, dout, ldo, dtable, n, D, V, p_drop,
              (unsigned long long)seed, (const unsigned long long*)seed_dev);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
