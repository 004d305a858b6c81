// Linear / Conv (channels-last implicit GEMM) forward and backward, plus the small elementwise
// helpers around them.  Replaces the nn.Linear / nn.Conv1d / nn.Conv2d call sites listed in
// include/s2ag.h.  All fp32, exact-order-independent up to fp32 summation order.
#include "s2ag.h"
#include "gemm.cuh"

using namespace s2ag;

#ifndef S2AG_EMU
namespace s2ag {
// umma_conv.cu: stride-1 convolutions as shifted-window tcgen05 contractions (weights stationary in shared memory)
bool conv_shift_launch(const float* x, long ldpix_x, int N, int H, int W, int Cin, const float* w, int w_mode, int w_ci,
                       const float* bias, float* y, long ldpix_y, int Cout, int KH, int KW, int ph, int pw, int Ho,
                       int Wo, int act, float slope, int accumulate, void* stream);
// umma_wgrad.cu: their weight gradients (shifted-window contraction over pixels, accumulators resident in TMEM)
bool conv_wgrad_shift_launch(const float* dy, long ldpix_dy, const float* x, long ldpix_x, int N, int H, int W, int Cin,
                             float* dw, int Cout, int KH, int KW, int ph, int pw, int Ho, int Wo, void* stream);
}
#endif

// ------------------------------------------------------------------ column sums (bias grads)
// db[n] += sum_m dy[m*ld + n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, long ld, float* __restrict__ db,
                                                     int M, int N, int rows_per_block) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int n = blockIdx.x * 32 + tx;
  const int mbeg = blockIdx.y * rows_per_block;
  int mend = mbeg + rows_per_block; if (mend > M) mend = M;
  float s = 0.f;
  if (n < N) for (int m = mbeg + ty; m < mend; m += 8) s += __ldg(dy + (long)m * ld + n);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(db + n, t);
  }
}
namespace s2ag {
void launch_colsum(const float* dy, long ld, float* db, int M, int N, void* stream) {
  int rpb = 256;
  dim3 grid(s2ag_cdiv(N, 32), s2ag_cdiv(M, rpb));
  auto kfn = &colsum_kernel;
  S2AG_LAUNCH(kfn, grid, 256, 0, stream, dy, ld, db, M, N, rpb);
}
}  // namespace s2ag

// ------------------------------------------------------------------ Linear
extern "C" int s2ag_linear_fwd(const float* x, long ldx, const float* w, const float* bias, float* y, long ldy,
                               int M, int N, int K, int act, float slope, void* stream) {
  S2AG_CHECK_ARG(x && w && y && M >= 0 && N > 0 && K > 0 && ldx >= K && ldy >= N);
  LdPlain<true> a{x, ldx, 1, 0};
  LdPlain<true> b{w, (long)K, 1, 0};
  launch_gemm(a, b, make_epi(y, ldy, bias, act, slope, 0), M, N, K, 1, 1, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_linear_bwd_data(const float* dy, long lddy, const float* w, float* dx, long lddx,
                                    int M, int N, int K, int accumulate, void* stream) {
  S2AG_CHECK_ARG(dy && w && dx && M >= 0 && N > 0 && K > 0 && lddy >= N && lddx >= K);
  // dx[m,k] = sum_n dy[m,n] w[n,k]  -> rows m, cols k, contraction n
  LdPlain<true> a{dy, lddy, 1, 0};
  LdPlain<false> b{w, 1, (long)K, 0};  // element(row=k, kk=n) = w[n*K + k]
  launch_gemm(a, b, make_epi(dx, lddx, nullptr, 0, 0.f, accumulate ? 1 : 0), M, K, N, 1, 1, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_linear_bwd_weight(const float* dy, long lddy, const float* x, long ldx, float* dw, float* db,
                                      int M, int N, int K, void* stream) {
  S2AG_CHECK_ARG(dy && x && dw && M >= 0 && N > 0 && K > 0 && lddy >= N && ldx >= K);
  if (M == 0) return S2AG_OK;
  // dw[n,k] += sum_m dy[m,n] x[m,k]  -> rows n, cols k, contraction m
  LdPlain<false> a{dy, 1, lddy, 0};
  LdPlain<false> b{x, 1, ldx, 0};
  int sk = pick_splitk(N, K, M, 1);
  launch_gemm(a, b, make_epi(dw, (long)K, nullptr, 0, 0.f, sk > 1 ? 2 : 1), N, K, M, 1, sk, stream);
  if (db) launch_colsum(dy, lddy, db, M, N, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

// ------------------------------------------------------------------ Linear over the row axis (batched, A transposed)
extern "C" int s2ag_linear_t_fwd(const float* x, const float* w, const float* bias, float* y, long ldy,
                                 int B, int L, int C, int N, int act, float slope, void* stream) {
  S2AG_CHECK_ARG(x && w && y && B >= 0 && L > 0 && C > 0 && N > 0 && ldy >= N);
  if (B == 0) return S2AG_OK;
  LdPlain<false> a{x, 1, (long)C, (long)L * C};  // element(row=c, k=l) = x[b][l][c]
  LdPlain<true> b{w, (long)L, 1, 0};
  EpiGeneric e = make_epi(y, ldy, bias, act, slope, 0);
  e.bstride = (long)C * ldy;
  launch_gemm(a, b, e, C, N, L, B, 1, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
extern "C" int s2ag_linear_t_bwd_data(const float* dy, long lddy, const float* w, float* dx,
                                      int B, int L, int C, int N, void* stream) {
  S2AG_CHECK_ARG(dy && w && dx && B >= 0 && L > 0 && C > 0 && N > 0 && lddy >= N);
  if (B == 0) return S2AG_OK;
  LdPlain<false> a{w, 1, (long)L, 0};                 // element(row=l, k=n) = w[n][l]
  LdPlain<true> b{dy, lddy, 1, (long)C * lddy};       // element(row=c, k=n) = dy[b][c][n]
  EpiGeneric e = make_epi(dx, (long)C);
  e.bstride = (long)L * C;
  launch_gemm(a, b, e, L, C, N, B, 1, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
extern "C" int s2ag_linear_t_bwd_weight(const float* dy, long lddy, const float* x, float* dw, float* db,
                                        int B, int L, int C, int N, void* stream) {
  S2AG_CHECK_ARG(dy && x && dw && B >= 0 && L > 0 && C > 0 && N > 0 && lddy >= N);
  if (B == 0) return S2AG_OK;
  LdPlain<false> a{dy, 1, lddy, (long)C * lddy};      // element(row=n, k=c) = dy[b][c][n]
  LdPlain<true> b{x, (long)C, 1, (long)L * C};        // element(row=l, k=c) = x[b][l][c]
  launch_gemm(a, b, make_epi(dw, (long)L, nullptr, 0, 0.f, 2), N, L, C, B, 1, stream);
  if (db) launch_colsum(dy, lddy, db, B * C, N, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

// ------------------------------------------------------------------ elementwise helpers
__global__ void act_bwd_kernel(const float* __restrict__ dy, long lddy, const float* __restrict__ y, long ldy,
                               float* __restrict__ dpre, long ldd, int M, int N, int act, float slope) {
  long total = (long)M * N;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int m = (int)(i / N), n = (int)(i % N);
    dpre[(long)m * ldd + n] = dy[(long)m * lddy + n] * s2ag_act_grad_from_out(y[(long)m * ldy + n], act, slope);
  }
}
extern "C" int s2ag_act_bwd(const float* dy, long lddy, const float* y, long ldy, float* dpre, long ldd,
                            int M, int N, int act, float slope, void* stream) {
  S2AG_CHECK_ARG(dy && y && dpre && M >= 0 && N > 0);
  long total = (long)M * N;
  if (total == 0) return S2AG_OK;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  auto kfn = &act_bwd_kernel;
  S2AG_LAUNCH(kfn, blocks, 256, 0, stream, dy, lddy, y, ldy, dpre, ldd, M, N, act, slope);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

__global__ void add_halves_kernel(const float* __restrict__ x, float* __restrict__ y, long M, int H) {
  long total = M * H;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long m = i / H; int h = (int)(i % H);
    y[i] = x[m * 2 * H + h] + x[m * 2 * H + H + h];
  }
}
extern "C" int s2ag_add_halves(const float* x, float* y, int M, int H, void* stream) {
  S2AG_CHECK_ARG(x && y && M >= 0 && H > 0);
  long total = (long)M * H;
  if (total == 0) return S2AG_OK;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  auto kfn = &add_halves_kernel;
  S2AG_LAUNCH(kfn, blocks, 256, 0, stream, x, y, (long)M, H);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long n, float p,
                               unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev) seed += seed_dev[0];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = x[i] * s2ag_dropout_scale(seed, (unsigned long long)i, p);
}
extern "C" int s2ag_dropout(const float* x, float* y, long n, float p, uint64_t seed, const uint64_t* seed_dev,
                            void* stream) {
  S2AG_CHECK_ARG(x && y && n >= 0 && p >= 0.f && p < 1.f);
  if (n == 0) return S2AG_OK;
  int blocks = (int)((n + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  auto kfn = &dropout_kernel;
  S2AG_LAUNCH(kfn, blocks, 256, 0, stream, x, y, n, p, (unsigned long long)seed, (const unsigned long long*)seed_dev);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

__global__ void seed_advance_kernel(unsigned long long* s, unsigned long long inc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) s[0] += inc;
}
extern "C" int s2ag_seed_advance(uint64_t* seed_dev, uint64_t inc, void* stream) {
  S2AG_CHECK_ARG(seed_dev);
  auto kfn = &seed_advance_kernel;
  S2AG_LAUNCH(kfn, 1, 32, 0, stream, (unsigned long long*)seed_dev, (unsigned long long)inc);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}


// ------------------------------------------------------------------ direct Conv1d for tiny Cin*K (WavEncoder conv1)
// y[n, lo, co] = act(b[co] + sum_{k,c} x[n, lo*s + k - p, c] * w[co, c, k]);  one thread = one output position, all
// Cout channels (weights in shared memory as [c*K + k][Cout]).  The implicit-GEMM engines waste their tiles on a
// contraction of length 15: this one is bound by writing y (16 channels x 4 bytes per position).
constexpr int DC_MAX_CK = 64, DC_MAX_CO = 32;
__global__ void __launch_bounds__(256) conv1d_direct_kernel(const float* __restrict__ x, long ldpix_x, int L, int Cin,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ y, long ldpix_y, int Lo, int Cout, int K,
                                                            int stride, int pad, long total, int act, float slope) {
  __shared__ float ws[DC_MAX_CK * DC_MAX_CO];
  __shared__ float bs[DC_MAX_CO];
  const int CK = Cin * K;
  for (int i = threadIdx.x; i < CK * Cout; i += blockDim.x) {
    const int co = i % Cout, ck = i / Cout;   // ck = c*K + k ; reference weight [co][c][k]
    ws[ck * Cout + co] = w[(long)co * CK + ck];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) bs[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  for (long pos = blockIdx.x * (long)blockDim.x + threadIdx.x; pos < total; pos += (long)gridDim.x * blockDim.x) {
    const int lo = (int)(pos % Lo); const long n = pos / Lo;
    float acc[DC_MAX_CO];
#pragma unroll
    for (int co = 0; co < DC_MAX_CO; ++co) acc[co] = co < Cout ? bs[co] : 0.f;
    const int l0 = lo * stride - pad;
    for (int k = 0; k < K; ++k) {
      const int l = l0 + k;
      if (l < 0 || l >= L) continue;
      const float* xp = x + (n * L + l) * ldpix_x;
      for (int c = 0; c < Cin; ++c) {
        const float xv = __ldg(xp + c);
        const float* wr = ws + (c * K + k) * Cout;
#pragma unroll
        for (int co = 0; co < DC_MAX_CO; ++co)
          if (co < Cout) acc[co] = fmaf(xv, wr[co], acc[co]);
      }
    }
    float* yp = y + pos * ldpix_y;
#pragma unroll
    for (int co = 0; co < DC_MAX_CO; ++co)
      if (co < Cout) yp[co] = s2ag_act(acc[co], act, slope);
  }
}

// ------------------------------------------------------------------ Conv (channels-last implicit GEMM)
static inline int conv_out(int L, int k, int s, int p, int d) { return (L + 2 * p - d * (k - 1) - 1) / s + 1; }

// ---- temporal-convolution weight gradient as a PLAIN contraction over a sliding-window view --------------------------
// For a stride-1 Conv1d over channels-last rows the im2col row of output (n, t) is the contiguous run
// x_pad[n, t .. t + Kt - 1, :] of the zero-padded input: with x_pad laid out [N, T + 2p, C] (+ Kt - 1 zero rows at the
// end) and dy_pad [N, T + 2p, Cout] (rows t >= T zero), the gradient is dwT[dt*C + c][co] = sum_r x_pad[(r + dt)*C + c] *
// dy_pad[r*Cout + co] over ALL R = N (T + 2p) rows: two row-major operands with pitches C and Cout, no gather, no clip
// boundaries (the windows of the dummy rows t >= T reach into the next clip and are multiplied by zero).  The composed
// ST-GCN convolution (stgcn.cu) is the user: 48 x 1296 x 17 408 through the transposed-im2col loader cost 140-370 us.
namespace {
__global__ void __launch_bounds__(256) pad_time_kernel(const float* __restrict__ src, long ld_src, float* __restrict__ dst,
                                                       int N, int T, int C, int Tp, int off, long total_rows) {
  const long total = total_rows * C;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long row = e / C; const int c = (int)(e - row * C);
    const long n = row / Tp; const int t = (int)(row - n * Tp) - off;
    dst[e] = (n < N && t >= 0 && t < T) ? __ldg(src + (n * T + t) * ld_src + c) : 0.f;
  }
}
}  // namespace

extern "C" int s2ag_pad_time(const float* src, long ld_src, float* dst, int N, int T, int C, int Tp, int off,
                             int tail_rows, void* stream) {
  S2AG_CHECK_ARG(src && dst && N >= 0 && T > 0 && C > 0 && Tp >= T + off && off >= 0 && tail_rows >= 0 && ld_src >= C);
  const long rows = (long)N * Tp + tail_rows;
  if (rows == 0) return S2AG_OK;
  long blocks = (rows * C + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
  auto kfn = &pad_time_kernel;
  S2AG_LAUNCH(kfn, (int)blocks, 256, 0, stream, src, ld_src, dst, N, T, C, Tp, off, rows);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_window_wgrad(const float* dy_pad, long lddy, const float* x_pad, long ldx, float* dwT, long R,
                                 int Cout, int Kwin, void* stream) {
  S2AG_CHECK_ARG(dy_pad && x_pad && dwT && R >= 0 && R < (1L << 31) && Cout > 0 && Kwin > 0 && lddy >= Cout && ldx > 0);
  if (R == 0) return S2AG_OK;
  // dwT[kcol, co] += sum_r x_pad[r*ldx + kcol] * dy_pad[r*lddy + co]: rows kcol (overlapping windows: ldx < Kwin), cols co
  LdPlain<false> a{x_pad, 1, ldx, 0};
  LdPlain<false> b{dy_pad, 1, lddy, 0};
  int sk = pick_splitk(Kwin, Cout, (int)R, 1);
  launch_gemm(a, b, make_epi(dwT, (long)Cout, nullptr, 0, 0.f, sk > 1 ? 2 : 1), Kwin, Cout, (int)R, 1, sk, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_colsum(const float* dy, long lddy, float* db, int M, int N, void* stream) {
  S2AG_CHECK_ARG(dy && db && M >= 0 && N > 0 && lddy >= N);
  if (M == 0) return S2AG_OK;
  launch_colsum(dy, lddy, db, M, N, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_conv_fwd(const float* x, long ldpix_x, int N, int H, int W, int Cin,
                             const float* w, const float* bias, float* y, long ldpix_y, int Cout,
                             int KH, int KW, int sh, int sw, int ph, int pw, int dh, int dw,
                             int act, float slope, void* stream) {
  S2AG_CHECK_ARG(x && w && y && N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0);
  S2AG_CHECK_ARG(sh > 0 && sw > 0 && dh > 0 && dw > 0 && ldpix_x >= Cin && ldpix_y >= Cout);
  int Ho = conv_out(H, KH, sh, ph, dh), Wo = conv_out(W, KW, sw, pw, dw);
  S2AG_CHECK_ARG(Ho > 0 && Wo > 0);
  if (W == 1 && KW == 1 && dh == 1 && Cin * KH <= DC_MAX_CK && Cout <= DC_MAX_CO && Cin * KH < 32) {
    // contraction shorter than one tensor-core k-block: direct kernel
    const long total = (long)N * Ho;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    auto kfn = &conv1d_direct_kernel;
    S2AG_LAUNCH(kfn, blocks, 256, 0, stream, x, ldpix_x, H, Cin, w, bias, y, ldpix_y, Ho, Cout, KH, sh, ph, total, act, slope);
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
#ifndef S2AG_EMU
  if (g_engine == 0 && sh == 1 && sw == 1 && dh == 1 && dw == 1 &&
      conv_shift_launch(x, ldpix_x, N, H, W, Cin, w, 0, Cin, bias, y, ldpix_y, Cout, KH, KW, ph, pw, Ho, Wo, act, slope, 0,
                        stream)) {
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
#endif
  // k order (kh, kw, c): the activation operand gathers contiguous channels of one pixel (16-byte loads when
  // Cin % 4 == 0); the weight operand reads the reference layout [Cout][Cin][KH][KW] through LdWkkc.
  LdConv<ORDER_KKC> a{x, H, W, Cin, Ho, Wo, KH, KW, sh, sw, dh, dw, +1, -ph, -pw, ldpix_x};
  int K = Cin * KH * KW;
  LdWkkc b{w, Cin, KH * KW};
  launch_gemm(a, b, make_epi(y, ldpix_y, bias, act, slope, 0), N * Ho * Wo, Cout, K, 1, 1, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_conv_bwd_data(const float* dy, long ldpix_dy, int N, int H, int W, int Cin,
                                  const float* w, float* dx, long ldpix_dx, int Cout,
                                  int KH, int KW, int ph, int pw, int dh, int dw, int accumulate, void* stream) {
  S2AG_CHECK_ARG(dy && w && dx && N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0);
  int Ho = conv_out(H, KH, 1, ph, dh), Wo = conv_out(W, KW, 1, pw, dw);
  S2AG_CHECK_ARG(Ho > 0 && Wo > 0 && ldpix_dy >= Cout && ldpix_dx >= Cin);
  // dx[n,hi,wi,c] = sum_{co,kh,kw} dy[n, hi+ph-kh*dh, wi+pw-kw*dw, co] * w[co,c,kh,kw]
#ifndef S2AG_EMU
  if (g_engine == 0 && dh == 1 && dw == 1 && KH - 1 - ph >= 0 && KW - 1 - pw >= 0 &&
      conv_shift_launch(dy, ldpix_dy, N, Ho, Wo, Cout, w, 1, Cin, nullptr, dx, ldpix_dx, Cin, KH, KW, KH - 1 - ph,
                        KW - 1 - pw, H, W, S2AG_ACT_NONE, 0.f, accumulate, stream)) {
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
#endif
  // contraction index (kh, kw, co): contiguous output channels of one dy pixel
  LdConv<ORDER_KKC> a{dy, Ho, Wo, Cout, H, W, KH, KW, 1, 1, dh, dw, -1, ph, pw, ldpix_dy};
  int KK = KH * KW;
  LdWdgrad<ORDER_KKC> b{w, Cout, KK, (long)Cin * KK, (long)KK, 1};
  launch_gemm(a, b, make_epi(dx, ldpix_dx, nullptr, 0, 0.f, accumulate ? 1 : 0), N * H * W, Cin, Cout * KK, 1, 1,
              stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_conv_bwd_weight(const float* dy, long ldpix_dy, const float* x, long ldpix_x,
                                    int N, int H, int W, int Cin, float* dw, float* db, int Cout,
                                    int KH, int KW, int sh, int sw, int ph, int pw, int dh, int dwd, void* stream) {
  S2AG_CHECK_ARG(dy && x && dw && N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0);
  int Ho = conv_out(H, KH, sh, ph, dh), Wo = conv_out(W, KW, sw, pw, dwd);
  S2AG_CHECK_ARG(Ho > 0 && Wo > 0 && ldpix_dy >= Cout && ldpix_x >= Cin);
  int Mrows = N * Ho * Wo;
  if (Mrows == 0) return S2AG_OK;
  int K = Cin * KH * KW;
#ifndef S2AG_EMU
  if (g_engine == 0 && sh == 1 && sw == 1 && dh == 1 && dwd == 1 &&
      conv_wgrad_shift_launch(dy, ldpix_dy, x, ldpix_x, N, H, W, Cin, dw, Cout, KH, KW, ph, pw, Ho, Wo, stream)) {
    if (db) launch_colsum(dy, ldpix_dy, db, Mrows, Cout, stream);
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
#endif
  // dw[co, kcol] += sum_row dy[row, co] * im2col(x)[row, kcol]
  LdPlain<false> a{dy, 1, ldpix_dy, 0};
  // columns in (kh, kw, c) order (consecutive columns = contiguous channels of x); the epilogue stores column
  // tap*Cin + c at the reference position c*KH*KW + tap of dw[co]
  LdT<LdConv<ORDER_KKC>> b{LdConv<ORDER_KKC>{x, H, W, Cin, Ho, Wo, KH, KW, sh, sw, dh, dwd, +1, -ph, -pw, ldpix_x}};
  int sk = pick_splitk(Cout, K, Mrows, 1);
  EpiGeneric e = make_epi(dw, (long)K, nullptr, 0, 0.f, sk > 1 ? 2 : 1);
  e.perm_C = Cin; e.perm_KK = KH * KW;
  launch_gemm(a, b, e, Cout, K, Mrows, 1, sk, stream);
  if (db) launch_colsum(dy, ldpix_dy, db, Mrows, Cout, stream);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
