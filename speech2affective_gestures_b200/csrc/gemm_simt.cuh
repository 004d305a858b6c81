// fp32 SIMT tiled contraction  C[m,n] = epi( sum_k A(m,k) * B(n,k) )  with pluggable operand
// loaders (plain strided, implicit-im2col over channels-last activations, weight views) and
// pluggable epilogues.  This is the exact-fp32 engine every dense op of the hot path can run
// on; the tcgen05 kernels (umma_*.cu) replace it shape by shape and are checked against it.
//
// Tile: 64x64x16, 256 threads, 4x4 register micro-tile, register-prefetch double buffering.
// grid = (tilesN, tilesM, nbatch*splitk).  split-K partial sums are combined by atomicAdd in the
// epilogue (only legal for linear epilogues: the weight-gradient calls).
#pragma once
#include "common.cuh"

namespace s2ag {

constexpr int GBM = 64, GBN = 64, GBK = 16, GTHREADS = 256;

// ---------------------------------------------------------------- loaders
// element(batch,row,k) = p[batch*bstride + row*ld_row + k*ld_k]
template <bool KCONTIG>
struct LdPlain {
  static constexpr bool kContig = KCONTIG;
  const float* p; long ld_row; long ld_k; long bstride;
  __device__ __forceinline__ float operator()(int b, int row, int k) const {
    return __ldg(p + b * bstride + (long)row * ld_row + (long)k * ld_k);
  }
  // 4 consecutive k (zero beyond kend); one 16-byte load when k is the contiguous axis and the address is aligned
  __device__ __forceinline__ float4 load4(int b, int row, int k, int kend) const {
    const float* q = p + b * bstride + (long)row * ld_row + (long)k * ld_k;
    if (KCONTIG && k + 4 <= kend && ld_k == 1 && (reinterpret_cast<unsigned long long>(q) & 15ull) == 0)
      return __ldg(reinterpret_cast<const float4*>(q));
    float4 r;
    r.x = k < kend ? __ldg(q) : 0.f;
    r.y = k + 1 < kend ? __ldg(q + ld_k) : 0.f;
    r.z = k + 2 < kend ? __ldg(q + 2 * ld_k) : 0.f;
    r.w = k + 3 < kend ? __ldg(q + 3 * ld_k) : 0.f;
    return r;
  }
};

// Implicit im2col over a channels-last activation x[N][H][W][C] (Conv1d: W == 1).
//   row -> (n, ho, wo);  k -> (c, kh, kw) [ORDER_CKK, PyTorch weight order] or (kh, kw, c) [ORDER_KKC]
//   source pixel: hi = ho*sh + sgn*kh*dh + off_h ; wi = wo*sw + sgn*kw*dw + off_w  (zero outside)
// forward conv: sgn=+1, off=-pad.  data-gradient of a stride-1 conv: run it over dY with sgn=-1, off=+pad.
constexpr int ORDER_CKK = 0, ORDER_KKC = 1;
template <int ORDER>
struct LdConv {
  static constexpr bool kContig = true;
  const float* x; int H, W, C;       // source geometry
  int Ho, Wo;                        // row space geometry
  int KH, KW, sh, sw, dh, dw, sgn, off_h, off_w;
  long ldpix;                        // floats between consecutive pixels of x (>= C; lets x be a column slice)
  __device__ __forceinline__ float operator()(int, int row, int k) const {
    int wo = row % Wo; int t = row / Wo; int ho = t % Ho; int n = t / Ho;
    int c, kh, kw;
    if (ORDER == ORDER_CKK) { kw = k % KW; int t2 = k / KW; kh = t2 % KH; c = t2 / KH; }
    else { c = k % C; int t2 = k / C; kw = t2 % KW; kh = t2 / KW; }
    int hi = ho * sh + sgn * kh * dh + off_h;
    int wi = wo * sw + sgn * kw * dw + off_w;
    if (hi < 0 || hi >= H || wi < 0 || wi >= W) return 0.f;
    return __ldg(x + ((long)(n * H + hi) * W + wi) * ldpix + c);
  }
  // 4 consecutive k: with the (kh, kw, c) order and C % 4 == 0 they are 4 channels of ONE pixel
  __device__ __forceinline__ float4 load4(int b, int row, int k, int kend) const {
    if (ORDER == ORDER_KKC && (C & 3) == 0 && k + 4 <= kend) {
      int wo = row % Wo; int t = row / Wo; int ho = t % Ho; int n = t / Ho;
      int c = k % C; int t2 = k / C; int kw = t2 % KW; int kh = t2 / KW;
      int hi = ho * sh + sgn * kh * dh + off_h;
      int wi = wo * sw + sgn * kw * dw + off_w;
      if (hi < 0 || hi >= H || wi < 0 || wi >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
      const float* q = x + ((long)(n * H + hi) * W + wi) * ldpix + c;
      if ((reinterpret_cast<unsigned long long>(q) & 15ull) == 0) return __ldg(reinterpret_cast<const float4*>(q));
      return make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
    }
    float4 r;
    r.x = k < kend ? (*this)(b, row, k) : 0.f;
    r.y = k + 1 < kend ? (*this)(b, row, k + 1) : 0.f;
    r.z = k + 2 < kend ? (*this)(b, row, k + 2) : 0.f;
    r.w = k + 3 < kend ? (*this)(b, row, k + 3) : 0.f;
    return r;
  }
};

// Transposed view of another loader (swap the roles of row and k).
template <class L>
struct LdT {
  static constexpr bool kContig = !L::kContig;
  L l;
  __device__ __forceinline__ float operator()(int b, int row, int k) const { return l(b, k, row); }
  __device__ __forceinline__ float4 load4(int b, int row, int k, int kend) const {
    float4 r;
    r.x = k < kend ? l(b, k, row) : 0.f;
    r.y = k + 1 < kend ? l(b, k + 1, row) : 0.f;
    r.z = k + 2 < kend ? l(b, k + 2, row) : 0.f;
    r.w = k + 3 < kend ? l(b, k + 3, row) : 0.f;
    return r;
  }
};

// Weight view for data-gradient GEMMs: element(row=c_in, k) with k -> (co, kk) [ORDER_CKK] or (kk, co) [ORDER_KKC]
// offset = co*s_co + c*s_c + kk*s_kk
template <int ORDER>
struct LdWdgrad {
  static constexpr bool kContig = false;  // consecutive k are strided in memory; consecutive rows (c_in) are the near axis
  const float* w; int Cout, KK; long s_co, s_c, s_kk;
  __device__ __forceinline__ float operator()(int, int row, int k) const {
    int co, kk;
    if (ORDER == ORDER_CKK) { kk = k % KK; co = k / KK; } else { co = k % Cout; kk = k / Cout; }
    return __ldg(w + co * s_co + row * s_c + kk * s_kk);
  }
  __device__ __forceinline__ float4 load4(int b, int row, int k, int kend) const {
    float4 r;
    r.x = k < kend ? (*this)(b, row, k) : 0.f;
    r.y = k + 1 < kend ? (*this)(b, row, k + 1) : 0.f;
    r.z = k + 2 < kend ? (*this)(b, row, k + 2) : 0.f;
    r.w = k + 3 < kend ? (*this)(b, row, k + 3) : 0.f;
    return r;
  }
};

// ---------------------------------------------------------------- epilogues
// v = alpha*acc (+bias[n]) ; v = act(v) ; v *= act'(mul_src[m,n]) ; store / add / atomicAdd
struct EpiGeneric {
  float* C; long ldc; long bstride;
  const float* bias; long bias_bstride;
  float alpha;
  int act; float slope;
  int mode;  // 0 store, 1 C += v, 2 atomicAdd
  const float* mul_src; long ld_mul; int mul_act; float mul_slope;
  __device__ __forceinline__ void operator()(int b, int m, int n, float acc, bool partial) const {
    float v = alpha * acc;
    if (bias) v += __ldg(bias + b * bias_bstride + n);
    v = s2ag_act(v, act, slope);
    if (mul_src) v *= s2ag_act_grad_from_out(__ldg(mul_src + (long)m * ld_mul + n), mul_act, mul_slope);
    float* dst = C + b * bstride + (long)m * ldc + n;
    if (partial || mode == 2) atomicAdd(dst, v);
    else if (mode == 1) *dst += v;
    else *dst = v;
  }
};
static inline EpiGeneric make_epi(float* C, long ldc, const float* bias = nullptr, int act = 0, float slope = 0.f,
                                  int mode = 0) {
  EpiGeneric e; e.C = C; e.ldc = ldc; e.bstride = 0; e.bias = bias; e.bias_bstride = 0; e.alpha = 1.f;
  e.act = act; e.slope = slope; e.mode = mode; e.mul_src = nullptr; e.ld_mul = 0; e.mul_act = 0; e.mul_slope = 0.f;
  return e;
}

// ---------------------------------------------------------------- kernel
template <class LdA, class LdB, class Epi>
__global__ void __launch_bounds__(GTHREADS) gemm_simt_kernel(LdA a, LdB b, Epi epi, int M, int N, int K, int splitk) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int n0 = blockIdx.x * GBN, m0 = blockIdx.y * GBM;
  const int batch = blockIdx.z / splitk, ks = blockIdx.z % splitk;
  int kper = (K + splitk - 1) / splitk;
  kper = ((kper + GBK - 1) / GBK) * GBK;
  const int kbeg = ks * kper;
  const int kend = (kbeg + kper < K) ? kbeg + kper : K;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * GTHREADS;
      int r, kk;
      if (LdA::kContig) { r = idx / GBK; kk = idx % GBK; } else { r = idx % GBM; kk = idx / GBM; }
      int m = m0 + r, k = k0 + kk;
      ra[i] = (m < M && k < kend) ? a(batch, m, k) : 0.f;
      if (LdB::kContig) { r = idx / GBK; kk = idx % GBK; } else { r = idx % GBN; kk = idx / GBN; }
      int n = n0 + r; k = k0 + kk;
      rb[i] = (n < N && k < kend) ? b(batch, n, k) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * GTHREADS;
      int r, kk;
      if (LdA::kContig) { r = idx / GBK; kk = idx % GBK; } else { r = idx % GBM; kk = idx / GBM; }
      As[kk][r] = ra[i];
      if (LdB::kContig) { r = idx / GBK; kk = idx % GBK; } else { r = idx % GBN; kk = idx / GBN; }
      Bs[kk][r] = rb[i];
    }
  };

  if (kbeg < kend) {
    fetch(kbeg);
    stash();
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += GBK) {
    const bool more = (k0 + GBK) < kend;
    if (more) fetch(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
    if (more) stash();
    __syncthreads();
  }
  if (kbeg < kend || splitk == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = m0 + ty * 4 + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = n0 + tx * 4 + j;
        if (n < N) epi(batch, m, n, acc[i][j], splitk > 1);
      }
    }
  }
}

// Host-side launcher.  splitk > 1 requires a linear epilogue (it will atomicAdd partial sums).
template <class LdA, class LdB, class Epi>
static inline void launch_gemm_simt(const LdA& a, const LdB& b, const Epi& epi, int M, int N, int K, int nbatch,
                                    int splitk, void* stream) {
  if (M <= 0 || N <= 0) return;
  if (splitk < 1) splitk = 1;
  auto kfn = &gemm_simt_kernel<LdA, LdB, Epi>;
  dim3 grid(s2ag_cdiv(N, GBN), s2ag_cdiv(M, GBM), nbatch * splitk);
  S2AG_LAUNCH(kfn, grid, GTHREADS, 0, stream, a, b, epi, M, N, K, splitk);
}

// split-K heuristic for weight-gradient contractions (few output tiles, very long K)
static inline int pick_splitk(int M, int N, int K, int nbatch) {
  long tiles = (long)s2ag_cdiv(M, GBM) * s2ag_cdiv(N, GBN) * nbatch;
  int want = (int)((148 * 2 + tiles - 1) / tiles);
  int maxk = K / 64; if (maxk < 1) maxk = 1;
  if (want > maxk) want = maxk;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return want;
}

}  // namespace s2ag
