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
// Every loader offers two interfaces:
//   float operator()(batch, row, k)                      -- random access (SIMT engine, tests)
//   Cur cursor(batch, row, k0); load8(cur, kend, v[8]); advance(cur)
//        -- streaming access for the tcgen05 engine: a cursor pins one operand row and walks k in steps of 32
//           (advance), load8 fetches the 8 consecutive k at the cursor (zeros beyond kend).  All row-dependent
//           index arithmetic (pixel decomposition, base pointers) is done once per tile in cursor(); mixed-radix
//           k indices are carried incrementally instead of being re-divided per element.
constexpr int LD_STEP = 32;  // k advance per cursor step (= the tcgen05 engine's k-block)

__device__ __forceinline__ bool s2ag_aligned16(const void* q) { return (reinterpret_cast<unsigned long long>(q) & 15ull) == 0; }

// element(batch,row,k) = p[batch*bstride + row*ld_row + k*ld_k]
template <bool KCONTIG>
struct LdPlain {
  static constexpr bool kContig = KCONTIG;
  const float* p; long ld_row; long ld_k; long bstride;
  __device__ __forceinline__ float operator()(int b, int row, int k) const {
    return __ldg(p + b * bstride + (long)row * ld_row + (long)k * ld_k);
  }
  struct Cur { const float* q; int k; int vec; };  // vec: 2 = one 32-byte load, 1 = two 16-byte loads, 0 = scalar
  __device__ __forceinline__ Cur cursor(int b, int row, int k) const {
    const float* q = p + b * bstride + (long)row * ld_row + (long)k * ld_k;
    int vec = 0;
    if (KCONTIG && ld_k == 1) {
      const unsigned long long a = reinterpret_cast<unsigned long long>(q);
      vec = (a & 31ull) == 0 ? 2 : ((a & 15ull) == 0 ? 1 : 0);  // advance() (+128 bytes) keeps the alignment
    }
    return Cur{q, k, vec};
  }
  __device__ __forceinline__ void load8_full(const Cur& c, float (&v)[8]) const {  // all 8 k in range
#ifndef S2AG_EMU
    if (c.vec == 2) {  // one full 32-byte sector per lane (LDG.E.256)
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                   : "l"(c.q));
      return;
    }
#endif
    if (c.vec) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(c.q)), b = __ldg(reinterpret_cast<const float4*>(c.q) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(c.q + i * ld_k);
    }
  }
  // FAST path of the tcgen05 engine: the host verified that EVERY row segment is 32-byte aligned (fast_ok), so an item
  // is exactly one LDG.E.256 with no per-thread alignment state
  static constexpr bool kHasFast = KCONTIG;
  bool fast_ok() const {
    return KCONTIG && ld_k == 1 && (reinterpret_cast<unsigned long long>(p) & 31ull) == 0 && (ld_row & 7) == 0 &&
           (bstride & 7) == 0;
  }
  __device__ __forceinline__ void load8_fast(const Cur& c, float (&v)[8]) const {
#ifndef S2AG_EMU
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(c.q));
#else
    load8_full(c, v);
#endif
  }
  __device__ __forceinline__ void load8(const Cur& c, int kend, float (&v)[8]) const {
    if (c.k + 8 <= kend) {
      load8_full(c, v);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (c.k + i < kend) ? __ldg(c.q + i * ld_k) : 0.f;
    }
  }
  __device__ __forceinline__ void advance(Cur& c) const { c.q += LD_STEP * ld_k; c.k += LD_STEP; }
};

// Implicit im2col over a channels-last activation x[N][H][W][C] (Conv1d: W == 1).
//   row -> (n, ho, wo);  k -> (c, kh, kw) [ORDER_CKK, PyTorch weight order] or (kh, kw, c) [ORDER_KKC]
//   source pixel: hi = ho*sh + sgn*kh*dh + off_h ; wi = wo*sw + sgn*kw*dw + off_w  (zero outside)
// forward conv: sgn=+1, off=-pad.  data-gradient of a stride-1 conv: run it over dY with sgn=-1, off=+pad.
constexpr int ORDER_CKK = 0, ORDER_KKC = 1;
template <int ORDER>
struct LdConv {
  static constexpr bool kContig = true;
  static constexpr int kOrder = ORDER;
  static constexpr bool kHasFast = false;
  bool fast_ok() const { return true; }
  template <class C> __device__ __forceinline__ void load8_fast(const C& c, float (&v)[8]) const { load8_full(c, v); }
  const float* x; int H, W, C;       // source geometry
  int Ho, Wo;                        // row space geometry
  int KH, KW, sh, sw, dh, dw, sgn, off_h, off_w;
  long ldpix;                        // floats between consecutive pixels of x (>= C; lets x be a column slice)
  __device__ __forceinline__ void split_k(int k, int& c, int& kh, int& kw) const {
    if (ORDER == ORDER_CKK) { kw = k % KW; int t2 = k / KW; kh = t2 % KH; c = t2 / KH; }
    else { c = k % C; int t2 = k / C; if (KW == 1) { kw = 0; kh = t2; } else { kw = t2 % KW; kh = t2 / KW; } }
  }
  __device__ __forceinline__ void split_row(int row, int& n, int& ho, int& wo) const {
    if (Wo == 1) { wo = 0; ho = row % Ho; n = row / Ho; } else { wo = row % Wo; int t = row / Wo; ho = t % Ho; n = t / Ho; }
  }
  __device__ __forceinline__ float operator()(int, int row, int k) const {
    int n, ho, wo, c, kh, kw;
    split_row(row, n, ho, wo);
    split_k(k, c, kh, kw);
    int hi = ho * sh + sgn * kh * dh + off_h;
    int wi = wo * sw + sgn * kw * dw + off_w;
    if (hi < 0 || hi >= H || wi < 0 || wi >= W) return 0.f;
    return __ldg(x + ((long)(n * H + hi) * W + wi) * ldpix + c);
  }
  struct Cur { const float* xn; int hi0, wi0, k, c, kh, kw; };
  __device__ __forceinline__ Cur cursor(int, int row, int k) const {
    int n, ho, wo;
    split_row(row, n, ho, wo);
    Cur cu;
    cu.xn = x + (long)n * H * W * ldpix;
    cu.hi0 = ho * sh + off_h; cu.wi0 = wo * sw + off_w; cu.k = k;
    split_k(k, cu.c, cu.kh, cu.kw);
    return cu;
  }
  __device__ __forceinline__ const float* pix(const Cur& cu, int kh, int kw) const {  // nullptr outside the image
    const int hi = cu.hi0 + sgn * kh * dh, wi = cu.wi0 + sgn * kw * dw;
    if (hi < 0 || hi >= H || wi < 0 || wi >= W) return nullptr;
    return cu.xn + ((long)hi * W + wi) * ldpix;
  }
  template <bool FULL>
  __device__ __forceinline__ void load8_impl(const Cur& cu, int kend, float (&v)[8]) const {
    int c = cu.c, kh = cu.kh, kw = cu.kw;
    if (ORDER == ORDER_KKC && (C & 3) == 0 && (FULL || cu.k + 8 <= kend)) {
      // two groups of 4 channels, each inside one pixel (k % 4 == 0 and C % 4 == 0)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float* q = pix(cu, kh, kw);
        if (q) {
          q += c;
          if (s2ag_aligned16(q)) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(q));
            v[4 * g] = a.x; v[4 * g + 1] = a.y; v[4 * g + 2] = a.z; v[4 * g + 3] = a.w;
          } else {
            v[4 * g] = __ldg(q); v[4 * g + 1] = __ldg(q + 1); v[4 * g + 2] = __ldg(q + 2); v[4 * g + 3] = __ldg(q + 3);
          }
        } else {
          v[4 * g] = v[4 * g + 1] = v[4 * g + 2] = v[4 * g + 3] = 0.f;
        }
        c += 4;
        if (c >= C) { c -= C; if (++kw == KW) { kw = 0; ++kh; } }
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float val = 0.f;
      if (FULL || cu.k + i < kend) {
        const float* q = pix(cu, kh, kw);
        if (q) val = __ldg(q + c);
      }
      v[i] = val;
      if (ORDER == ORDER_CKK) { if (++kw == KW) { kw = 0; if (++kh == KH) { kh = 0; ++c; } } }
      else { if (++c == C) { c = 0; if (++kw == KW) { kw = 0; ++kh; } } }
    }
  }
  __device__ __forceinline__ void load8(const Cur& cu, int kend, float (&v)[8]) const { load8_impl<false>(cu, kend, v); }
  __device__ __forceinline__ void load8_full(const Cur& cu, float (&v)[8]) const { load8_impl<true>(cu, 0, v); }
  __device__ __forceinline__ void advance(Cur& cu) const {
    cu.k += LD_STEP;
    if (ORDER == ORDER_KKC) {
      cu.c += LD_STEP;
      while (cu.c >= C) { cu.c -= C; if (++cu.kw == KW) { cu.kw = 0; ++cu.kh; } }
    } else {
      split_k(cu.k, cu.c, cu.kh, cu.kw);
    }
  }
};

// Transposed view of an implicit-im2col loader (swap the roles of row and k): element(row, k) = l(k, row).
// row = the conv's k index (c, kh, kw), fixed per cursor; k walks the pixels (n, ho, wo).
template <class L>
struct LdT {
  static constexpr bool kContig = !L::kContig;
  static constexpr bool kHasFast = false;
  bool fast_ok() const { return true; }
  template <class C> __device__ __forceinline__ void load8_fast(const C& c, float (&v)[8]) const { load8_full(c, v); }
  L l;
  __device__ __forceinline__ float operator()(int b, int row, int k) const { return l(b, k, row); }
  struct Cur { int c, kh, kw, p, n, ho, wo; };
  __device__ __forceinline__ Cur cursor(int, int row, int k) const {
    Cur cu;
    l.split_k(row, cu.c, cu.kh, cu.kw);
    cu.p = k;
    l.split_row(k, cu.n, cu.ho, cu.wo);
    return cu;
  }
  template <bool FULL>
  __device__ __forceinline__ void load8_impl(const Cur& cu, int kend, float (&v)[8]) const {
    int n = cu.n, ho = cu.ho, wo = cu.wo;
    const int dhh = l.sgn * cu.kh * l.dh + l.off_h, dww = l.sgn * cu.kw * l.dw + l.off_w;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float val = 0.f;
      if (FULL || cu.p + i < kend) {
        const int hi = ho * l.sh + dhh, wi = wo * l.sw + dww;
        if (hi >= 0 && hi < l.H && wi >= 0 && wi < l.W) val = __ldg(l.x + ((long)(n * l.H + hi) * l.W + wi) * l.ldpix + cu.c);
      }
      v[i] = val;
      if (++wo == l.Wo) { wo = 0; if (++ho == l.Ho) { ho = 0; ++n; } }
    }
  }
  __device__ __forceinline__ void load8(const Cur& cu, int kend, float (&v)[8]) const { load8_impl<false>(cu, kend, v); }
  __device__ __forceinline__ void load8_full(const Cur& cu, float (&v)[8]) const { load8_impl<true>(cu, 0, v); }
  __device__ __forceinline__ void advance(Cur& cu) const {
    cu.p += LD_STEP;
    l.split_row(cu.p, cu.n, cu.ho, cu.wo);
  }
};

// Weight view for data-gradient GEMMs: element(row=c_in, k) with k -> (co, kk) [ORDER_CKK] or (kk, co) [ORDER_KKC]
// offset = co*s_co + c*s_c + kk*s_kk
template <int ORDER>
struct LdWdgrad {
  static constexpr bool kContig = false;  // consecutive k are strided in memory; consecutive rows (c_in) are the near axis
  static constexpr bool kHasFast = false;
  bool fast_ok() const { return true; }
  template <class C> __device__ __forceinline__ void load8_fast(const C& c, float (&v)[8]) const { load8_full(c, v); }
  const float* w; int Cout, KK; long s_co, s_c, s_kk;
  __device__ __forceinline__ void split_k(int k, int& co, int& kk) const {
    if (ORDER == ORDER_CKK) { kk = k % KK; co = k / KK; } else { co = k % Cout; kk = k / Cout; }
  }
  __device__ __forceinline__ float operator()(int, int row, int k) const {
    int co, kk;
    split_k(k, co, kk);
    return __ldg(w + co * s_co + row * s_c + kk * s_kk);
  }
  struct Cur { const float* wr; int k, co, kk; };
  __device__ __forceinline__ Cur cursor(int, int row, int k) const {
    Cur cu;
    cu.wr = w + row * s_c; cu.k = k;
    split_k(k, cu.co, cu.kk);
    return cu;
  }
  template <bool FULL>
  __device__ __forceinline__ void load8_impl(const Cur& cu, int kend, float (&v)[8]) const {
    int co = cu.co, kk = cu.kk;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = (FULL || cu.k + i < kend) ? __ldg(cu.wr + co * s_co + kk * s_kk) : 0.f;
      if (ORDER == ORDER_CKK) { if (++kk == KK) { kk = 0; ++co; } } else { if (++co == Cout) { co = 0; ++kk; } }
    }
  }
  __device__ __forceinline__ void load8(const Cur& cu, int kend, float (&v)[8]) const { load8_impl<false>(cu, kend, v); }
  __device__ __forceinline__ void load8_full(const Cur& cu, float (&v)[8]) const { load8_impl<true>(cu, 0, v); }
  __device__ __forceinline__ void advance(Cur& cu) const {
    cu.k += LD_STEP;
    split_k(cu.k, cu.co, cu.kk);
  }
};

// Convolution weight in the reference layout w[Cout][Cin][KK] read with the k order (tap, c) of LdConv<ORDER_KKC>:
// element(row = co, k = tap*Cin + c) = w[(co*Cin + c)*KK + tap].  (The activation operand then gathers contiguous
// channels; the weight tensor is small and cache-resident, its strided reads are cheap.)
struct LdWkkc {
  static constexpr bool kContig = true;
  static constexpr bool kHasFast = false;
  bool fast_ok() const { return true; }
  template <class C> __device__ __forceinline__ void load8_fast(const C& c, float (&v)[8]) const { load8_full(c, v); }
  const float* w; int Cin, KK;
  __device__ __forceinline__ float operator()(int, int row, int k) const {
    const int c = k % Cin, tap = k / Cin;
    return __ldg(w + ((long)row * Cin + c) * KK + tap);
  }
  struct Cur { const float* wr; int k, c, tap; };
  __device__ __forceinline__ Cur cursor(int, int row, int k) const {
    return Cur{w + (long)row * Cin * KK, k, k % Cin, k / Cin};
  }
  template <bool FULL>
  __device__ __forceinline__ void load8_impl(const Cur& cu, int kend, float (&v)[8]) const {
    int c = cu.c, tap = cu.tap;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = (FULL || cu.k + i < kend) ? __ldg(cu.wr + (long)c * KK + tap) : 0.f;
      if (++c == Cin) { c = 0; ++tap; }
    }
  }
  __device__ __forceinline__ void load8(const Cur& cu, int kend, float (&v)[8]) const { load8_impl<false>(cu, kend, v); }
  __device__ __forceinline__ void load8_full(const Cur& cu, float (&v)[8]) const { load8_impl<true>(cu, 0, v); }
  __device__ __forceinline__ void advance(Cur& cu) const {
    cu.k += LD_STEP; cu.c += LD_STEP;
    while (cu.c >= Cin) { cu.c -= Cin; ++cu.tap; }
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
  int perm_C, perm_KK;  // perm_C > 0: output column n = tap*perm_C + c is stored at column c*perm_KK + tap
  __device__ __forceinline__ void operator()(int b, int m, int n, float acc, bool partial) const {
    float v = alpha * acc;
    if (bias) v += __ldg(bias + b * bias_bstride + n);
    v = s2ag_act(v, act, slope);
    if (mul_src) v *= s2ag_act_grad_from_out(__ldg(mul_src + (long)m * ld_mul + n), mul_act, mul_slope);
    const int ncol = perm_C > 0 ? (n % perm_C) * perm_KK + n / perm_C : n;
    float* dst = C + b * bstride + (long)m * ldc + ncol;
    if (partial || mode == 2) atomicAdd(dst, v);
    else if (mode == 1) *dst += v;
    else *dst = v;
  }
  // Column-major walk used by the tcgen05 epilogue (a lane owns one output column n and visits 32 rows): everything
  // that depends only on (batch, n) -- bias value, permuted column, base pointers -- is computed once in col().
  struct Col { float* dst; const float* mul; float bias; };
  __device__ __forceinline__ Col col(int b, int n) const {
    Col c;
    const int ncol = perm_C > 0 ? (n % perm_C) * perm_KK + n / perm_C : n;
    c.dst = C + b * bstride + ncol;
    c.mul = mul_src ? mul_src + n : nullptr;
    c.bias = bias ? __ldg(bias + b * bias_bstride + n) : 0.f;
    return c;
  }
  __device__ __forceinline__ void apply(const Col& c, int m, float acc, bool partial) const {
    float v = fmaf(alpha, acc, c.bias);
    v = s2ag_act(v, act, slope);
    if (c.mul) v *= s2ag_act_grad_from_out(__ldg(c.mul + (long)m * ld_mul), mul_act, mul_slope);
    float* dst = c.dst + (long)m * ldc;
    if (partial || mode == 2) atomicAdd(dst, v);
    else if (mode == 1) *dst += v;
    else *dst = v;
  }
  // Row-vector walk (a thread owns one output row and visits 8 consecutive columns at a time: two 16-byte stores, no
  // transposition through shared memory): the common store / accumulate epilogues without column permutation or mask.
  static constexpr bool kRowVec = true;
  __device__ __forceinline__ bool row_vec_ok(bool partial) const {
    return !partial && mode != 2 && perm_C == 0 && mul_src == nullptr && (ldc & 3) == 0 && (bstride & 3) == 0 &&
           (reinterpret_cast<unsigned long long>(C) & 15ull) == 0;
  }
  __device__ __forceinline__ const float* bias_of(int b) const { return bias ? bias + b * bias_bstride : nullptr; }
  // columns n .. n+7 (n % 4 == 0, all < N) of row m; bias8: this column group's bias values (zeros if there is no bias)
  __device__ __forceinline__ void store_row8(int b, int m, int n, const float (&acc)[8], const float (&bias8)[8]) const {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = s2ag_act(fmaf(alpha, acc[i], bias8[i]), act, slope);
    float4* dst = reinterpret_cast<float4*>(C + b * bstride + (long)m * ldc + n);
    if (mode == 1) {
      const float4 o0 = dst[0], o1 = dst[1];
      v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w; v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
    }
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
// does an epilogue functor offer the row-vector path (store_row8) of the tcgen05 engines?
template <class E, class = void> struct EpiHasRowVec { static constexpr bool value = false; };
template <class E> struct EpiHasRowVec<E, decltype((void)E::kRowVec)> { static constexpr bool value = E::kRowVec; };

static inline EpiGeneric make_epi(float* C, long ldc, const float* bias = nullptr, int act = 0, float slope = 0.f,
                                  int mode = 0) {
  EpiGeneric e; e.C = C; e.ldc = ldc; e.bstride = 0; e.bias = bias; e.bias_bstride = 0; e.alpha = 1.f;
  e.act = act; e.slope = slope; e.mode = mode; e.mul_src = nullptr; e.ld_mul = 0; e.mul_act = 0; e.mul_slope = 0.f;
  e.perm_C = 0; e.perm_KK = 0;
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
