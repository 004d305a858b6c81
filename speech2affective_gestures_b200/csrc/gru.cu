// One bidirectional nn.GRU layer, forward and BPTT backward, channels-last [B,T,*].
// Reference call sites: net/multimodal_context_net_v2.py:480-481,541 (G), :281-282,333 (frozen
// tri-modal baseline), :558-560,576 (D).  PyTorch semantics: gate order r,z,n;
//   r = s(Wir x + bir + Whr h + bhr); z = s(...); n = tanh(Win x + bin + r*(Whn h + bhn));
//   h' = (1-z)*n + z*h ; h0 = 0 ; output [fwd | rev].
// Structure: one time-batched input projection (both directions in one batched GEMM), then T
// recurrent launches, each fusing h @ W_hh^T for the three gates with the gate math (both
// directions in the same launch, grid.z).  Backward mirrors it: per step a gate-gradient kernel
// and a dh += dGh @ W_hh GEMM, then four time-batched weight-gradient GEMMs.
#include "s2ag.h"
#include "gemm.cuh"

using namespace s2ag;
namespace s2ag {
void launch_colsum(const float* dy, long ld, float* db, int M, int N, void* stream);
#ifndef S2AG_EMU
// umma_gru.cu: persistent weight-stationary recurrence on tcgen05
bool gru_persist_supported(int H);
// umma_gru_cluster.cu: recurrence as thread-block clusters exchanging h through distributed shared memory
bool gru_cluster_supported(int H);
int gru_cluster_fwd(const float* gi, const float* whh_f, long whh_dstride, const float* bhh_f, long bhh_dstride, float* out,
                    float* gates, int B, int T, int H, int x3, void* stream);
bool gru_cluster_fwd_selected(int H, int B);
bool gru_cluster_bwd_selected(int H, int B);
int gru_cluster_bwd(const float* dout, long lddout, int dir_stride, const float* out, const float* gates,
                    const float* whh_f, long whh_dstride, float* dgi, float* dgh, int B, int T, int H, int x3,
                    void* stream);
size_t gru_persist_ws_bytes(int B, int H);
int gru_persist_fwd(const float* gi, const float* whh_f, long whh_dstride, const float* bhh_f, long bhh_dstride,
                    float* out, float* gates, void* ws, int B, int T, int H, int x3, void* stream);
size_t gru_persist_bwd_ws_bytes(int B, int H);
int gru_persist_bwd(const float* dout, long lddout, int dir_stride, const float* out, const float* gates,
                    const float* whh_f, long whh_dstride, float* dgi, float* dgh, void* ws, int B, int T, int H, int x3,
                    void* stream);
#endif
}  // namespace s2ag

namespace {

constexpr int RB = 32, RJ = 32, RK = 16;  // batch rows, hidden units, k-chunk per CTA

// grid (ceil(H/RJ), ceil(B/RB), 2)
__global__ void __launch_bounds__(256) gru_step_kernel(
    const float* __restrict__ gi, const float* __restrict__ whh_f, long whh_dstride, const float* __restrict__ bhh_f,
    long bhh_dstride, float* __restrict__ out, float* __restrict__ gates, int B, int T, int H, int step) {
  __shared__ float As[RK][RB + 1];
  __shared__ float Ws[3][RK][RJ + 1];
  const int dir = blockIdx.z;
  const int t = dir == 0 ? step : T - 1 - step;
  const int tprev = dir == 0 ? t - 1 : t + 1;
  const bool has_prev = step > 0;
  const float* whh = whh_f + dir * whh_dstride;
  const float* bhh = bhh_f + dir * bhh_dstride;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int j0 = blockIdx.x * RJ, b0 = blockIdx.y * RB;
  float acc[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = 0.f; }
  if (has_prev) {
    for (int k0 = 0; k0 < H; k0 += RK) {
      // h_prev tile: rows b0.., k contiguous
      for (int idx = threadIdx.x; idx < RB * RK; idx += 256) {
        const int r = idx / RK, kk = idx % RK;
        const int b = b0 + r, k = k0 + kk;
        As[kk][r] = (b < B && k < H) ? out[((long)b * T + tprev) * 2 * H + dir * H + k] : 0.f;
      }
      for (int idx = threadIdx.x; idx < 3 * RJ * RK; idx += 256) {
        const int kk = idx % RK; const int jj = (idx / RK) % RJ; const int g = idx / (RK * RJ);
        const int j = j0 + jj, k = k0 + kk;
        Ws[g][kk][jj] = (j < H && k < H) ? __ldg(whh + ((long)g * H + j) * H + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < RK; ++kk) {
        const float w0 = Ws[0][kk][tx], w1 = Ws[1][kk][tx], w2 = Ws[2][kk][tx];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = As[kk][ty * 4 + i];
          acc[i][0] = fmaf(a, w0, acc[i][0]);
          acc[i][1] = fmaf(a, w1, acc[i][1]);
          acc[i][2] = fmaf(a, w2, acc[i][2]);
        }
      }
      __syncthreads();
    }
  }
  const int j = j0 + tx;
  if (j < H) {
    const float br = bhh[j], bz = bhh[H + j], bn = bhh[2 * H + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + ty * 4 + i;
      if (b >= B) continue;
      const long row = (long)b * T + t;
      const float* gir = gi + row * 6 * H + dir * 3 * H;
      const float ghn = acc[i][2] + bn;
      const float r = s2ag_sigmoid(gir[j] + acc[i][0] + br);
      const float z = s2ag_sigmoid(gir[H + j] + acc[i][1] + bz);
      const float n = tanhf(gir[2 * H + j] + r * ghn);
      const float hp = has_prev ? out[((long)b * T + tprev) * 2 * H + dir * H + j] : 0.f;
      out[row * 2 * H + dir * H + j] = (1.f - z) * n + z * hp;
      if (gates) {  // saved gates: [t][dir][gate][j][b]
        float* gs = gates + ((((long)t * 2 + dir) * 4) * H + j) * B + b;
        const long gstride = (long)H * B;
        gs[0] = r; gs[gstride] = z; gs[2 * gstride] = n; gs[3 * gstride] = ghn;
      }
    }
  }
}

// per step: gate gradients.  grid-stride over (dir, b, j).
// dh = dout[b,t,dir] + carry_in[dir][b][j];  writes dgi, dgh at (b,t,dir) and carry_out = dh*z
__global__ void gru_bwd_gate_kernel(const float* __restrict__ dout, long lddout, int dir_stride,
                                    const float* __restrict__ out, const float* __restrict__ gates,
                                    const float* __restrict__ carry_in, float* __restrict__ carry_out,
                                    float* __restrict__ dgi, float* __restrict__ dgh, int B, int T, int H, int step) {
  const long total = 2L * B * H;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i % H); const int b = (int)((i / H) % B); const int dir = (int)(i / ((long)H * B));
    // backward visits the forward order reversed: forward step index fs = T-1-step
    const int fs = T - 1 - step;
    const int t = dir == 0 ? fs : T - 1 - fs;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const long row = (long)b * T + t;
    float dh = dout[row * lddout + (long)dir * dir_stride + j];
    if (step > 0) dh += carry_in[((long)dir * B + b) * H + j];
    const float* gs = gates + ((((long)t * 2 + dir) * 4) * H + j) * B + b;  // [t][dir][gate][j][b]
    const long gstride = (long)H * B;
    const float r = gs[0], z = gs[gstride], n = gs[2 * gstride], ghn = gs[3 * gstride];
    const float hp = fs > 0 ? out[((long)b * T + tprev) * 2 * H + dir * H + j] : 0.f;
    const float dn = dh * (1.f - z) * (1.f - n * n);
    const float dz = dh * (hp - n) * z * (1.f - z);
    const float dr = dn * ghn * r * (1.f - r);
    float* a = dgi + (row * 2 + dir) * 3 * H;
    float* c = dgh + (row * 2 + dir) * 3 * H;
    a[j] = dr; a[H + j] = dz; a[2 * H + j] = dn;
    c[j] = dr; c[H + j] = dz; c[2 * H + j] = dn * r;
    carry_out[((long)dir * B + b) * H + j] = dh * z;
  }
}

}  // namespace

extern "C" long s2ag_gru_fwd_ws_floats(int B, int T, int H) {
  long n = (long)B * T * 6 * H;
#ifndef S2AG_EMU
  n += (long)((gru_persist_ws_bytes(B, H) + 3) / 4);
#endif
  return n;
}

extern "C" long s2ag_gru_bwd_ws_floats(int B, int T, int H) {
  long n = 12L * B * T * H + 4L * B * H;
#ifndef S2AG_EMU
  n += (long)((gru_persist_bwd_ws_bytes(B, H) + 3) / 4);
#endif
  return n;
}

extern "C" int s2ag_gru_layer_fwd(const float* x, long ldx, const float* w_ih_f, const float* w_ih_r,
                                  const float* b_ih_f, const float* b_ih_r, const float* w_hh_f, const float* w_hh_r,
                                  const float* b_hh_f, const float* b_hh_r, float* gi_ws, float* out, float* gates,
                                  int B, int T, int In, int H, void* stream) {
  S2AG_CHECK_ARG(x && w_ih_f && w_ih_r && b_ih_f && b_ih_r && w_hh_f && w_hh_r && b_hh_f && b_hh_r && gi_ws && out);
  S2AG_CHECK_ARG(B >= 0 && T > 0 && In > 0 && H > 0 && ldx >= In);
  if (B == 0) return S2AG_OK;
  const int M = B * T;
  {  // gi[m][dir][3H] = x @ W_ih[dir]^T + b_ih[dir]
    LdPlain<true> a{x, ldx, 1, 0};
    LdPlain<true> b{w_ih_f, (long)In, 1, (long)(w_ih_r - w_ih_f)};
    EpiGeneric e = make_epi(gi_ws, 6L * H, b_ih_f);
    e.bstride = 3L * H; e.bias_bstride = (long)(b_ih_r - b_ih_f);
    launch_gemm(a, b, e, M, 3 * H, In, 2, 1, stream);
  }
#ifndef S2AG_EMU
  if (g_engine == 0 && gru_cluster_fwd_selected(H, B)) {
    // one launch for all T steps of both directions: clusters of slice CTAs exchanging h through DSMEM
    int rc = gru_cluster_fwd(gi_ws, w_hh_f, (long)(w_hh_r - w_hh_f), b_hh_f, (long)(b_hh_r - b_hh_f), out, gates, B, T, H,
                             umma::g_precision == 0 ? 1 : 0, stream);
    if (rc != S2AG_OK) { s2ag_set_error("gru_cluster_fwd failed (%d)", rc); return rc; }
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
  if (g_engine == 0 && gru_persist_supported(H) && gru_persist_ws_bytes(B, H) > 0) {
    // one persistent launch (per <= 148-CTA batch chunk) for all T steps of both directions
    int rc = gru_persist_fwd(gi_ws, w_hh_f, (long)(w_hh_r - w_hh_f), b_hh_f, (long)(b_hh_r - b_hh_f), out, gates,
                             gi_ws + (long)M * 6 * H, B, T, H, umma::g_precision == 0 ? 1 : 0, stream);
    if (rc != S2AG_OK) { s2ag_set_error("gru_persist_fwd failed (%d)", rc); return rc; }
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
#endif
  dim3 grid(s2ag_cdiv(H, RJ), s2ag_cdiv(B, RB), 2);
  auto kfn = &gru_step_kernel;
  for (int s = 0; s < T; ++s)
    S2AG_LAUNCH(kfn, grid, 256, 0, stream, (const float*)gi_ws, w_hh_f, (long)(w_hh_r - w_hh_f), b_hh_f,
                (long)(b_hh_r - b_hh_f), out, gates, B, T, H, s);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

// The recurrence alone (gi precomputed): the latency-bound "GRU step" kernel, exposed for measurement (bench.py roofline)
// and for callers that batch the input projection themselves.
extern "C" int s2ag_gru_recurrence_fwd(const float* gi_ws, const float* w_hh_f, const float* w_hh_r, const float* b_hh_f,
                                       const float* b_hh_r, float* out, float* gates, int B, int T, int H, void* stream) {
  S2AG_CHECK_ARG(gi_ws && w_hh_f && w_hh_r && b_hh_f && b_hh_r && out && B > 0 && T > 0 && H > 0);
#ifndef S2AG_EMU
  if (g_engine == 0 && gru_cluster_fwd_selected(H, B)) {
    int rc = gru_cluster_fwd(gi_ws, w_hh_f, (long)(w_hh_r - w_hh_f), b_hh_f, (long)(b_hh_r - b_hh_f), out, gates, B, T, H,
                             umma::g_precision == 0 ? 1 : 0, stream);
    if (rc != S2AG_OK) { s2ag_set_error("gru_cluster_fwd failed (%d)", rc); return rc; }
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
  if (g_engine == 0 && gru_persist_supported(H) && gru_persist_ws_bytes(B, H) > 0) {
    int rc = gru_persist_fwd(gi_ws, w_hh_f, (long)(w_hh_r - w_hh_f), b_hh_f, (long)(b_hh_r - b_hh_f), out, gates,
                             const_cast<float*>(gi_ws) + (long)B * T * 6 * H, B, T, H, umma::g_precision == 0 ? 1 : 0,
                             stream);
    if (rc != S2AG_OK) { s2ag_set_error("gru_persist_fwd failed (%d)", rc); return rc; }
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
#endif
  dim3 grid(s2ag_cdiv(H, RJ), s2ag_cdiv(B, RB), 2);
  auto kfn = &gru_step_kernel;
  for (int s = 0; s < T; ++s)
    S2AG_LAUNCH(kfn, grid, 256, 0, stream, gi_ws, w_hh_f, (long)(w_hh_r - w_hh_f), b_hh_f, (long)(b_hh_r - b_hh_f), out,
                gates, B, T, H, s);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_gru_layer_bwd(const float* dout, long lddout, int dir_stride, const float* x, long ldx,
                                  const float* out, const float* gates,
                                  const float* w_ih_f, const float* w_ih_r, const float* w_hh_f, const float* w_hh_r,
                                  float* dx, long lddx, float* dw_ih_f, float* dw_ih_r, float* db_ih_f, float* db_ih_r,
                                  float* dw_hh_f, float* dw_hh_r, float* db_hh_f, float* db_hh_r, float* ws,
                                  int B, int T, int In, int H, int phases, void* stream) {
  S2AG_CHECK_ARG(dout && x && out && gates && w_ih_f && w_ih_r && w_hh_f && w_hh_r && ws);
  S2AG_CHECK_ARG(phases > 0 && phases < 8);
  const bool do_rec = phases & 1, do_dx = phases & 2, do_w = phases & 4;
  const bool want_w = dw_ih_f != nullptr;  // all eight weight-gradient pointers or none
  S2AG_CHECK_ARG(!want_w || (dw_ih_r && db_ih_f && db_ih_r && dw_hh_f && dw_hh_r && db_hh_f && db_hh_r));
  S2AG_CHECK_ARG(B >= 0 && T > 0 && In > 0 && H > 0 && ldx >= In);
  if (B == 0) return S2AG_OK;
  const int M = B * T;
  float* dgi = ws;
  float* dgh = ws + (long)M * 6 * H;
  float* carry[2] = {dgh + (long)M * 6 * H, dgh + (long)M * 6 * H + 2L * B * H};
  const long total = 2L * B * H;
  int eblocks = (int)((total + 255) / 256); if (eblocks > 148 * 8) eblocks = 148 * 8;
  auto kg = &gru_bwd_gate_kernel;
  bool persistent = false;
#ifndef S2AG_EMU
  if (do_rec && g_engine == 0 && gru_cluster_bwd_selected(H, B)) {
    int rc = gru_cluster_bwd(dout, lddout, dir_stride, out, gates, w_hh_f, (long)(w_hh_r - w_hh_f), dgi, dgh, B, T, H,
                             umma::g_precision == 0 ? 1 : 0, stream);
    if (rc != S2AG_OK) { s2ag_set_error("gru_cluster_bwd failed (%d)", rc); return rc; }
    persistent = true;
  } else if (do_rec && g_engine == 0 && gru_persist_bwd_ws_bytes(B, H) > 0) {
    int rc = gru_persist_bwd(dout, lddout, dir_stride, out, gates, w_hh_f, (long)(w_hh_r - w_hh_f), dgi, dgh,
                             carry[0] + 4L * B * H, B, T, H, umma::g_precision == 0 ? 1 : 0, stream);
    if (rc != S2AG_OK) { s2ag_set_error("gru_persist_bwd failed (%d)", rc); return rc; }
    persistent = true;
  }
#endif
  for (int s = 0; s < T && !persistent && do_rec; ++s) {
    float* cin = carry[s & 1];
    float* cout = carry[(s + 1) & 1];
    S2AG_LAUNCH(kg, eblocks, 256, 0, stream, dout, lddout, dir_stride, out, gates, (const float*)cin, cout, dgi, dgh,
                B, T, H, s);
    if (s + 1 < T) {
      // cout[dir][b][k] += sum_c dgh[b, t_dir, dir, c] * W_hh[dir][c, k]
      const int fs = T - 1 - s;
      const int t0 = fs, t1 = T - 1 - fs;
      LdPlain<true> a{dgh + (long)(t0 * 2 + 0) * 3 * H, (long)T * 6 * H, 1, (long)((t1 - t0) * 2 + 1) * 3 * H};
      LdPlain<false> b{w_hh_f, 1, (long)H, (long)(w_hh_r - w_hh_f)};
      EpiGeneric e = make_epi(cout, (long)H, nullptr, 0, 0.f, 1);
      e.bstride = (long)B * H;
      launch_gemm(a, b, e, B, H, 3 * H, 2, 1, stream);
    }
  }
  // weight / bias gradients, time-batched
  const float* w_ih[2] = {w_ih_f, w_ih_r};
  float* dw_ih[2] = {dw_ih_f, dw_ih_r};
  float* db_ih[2] = {db_ih_f, db_ih_r};
  float* dw_hh[2] = {dw_hh_f, dw_hh_r};
  float* db_hh[2] = {db_hh_f, db_hh_r};
  for (int d = 0; d < 2 && want_w && do_w; ++d) {
    {  // dW_ih[d][3H, In] += dgi_d^T @ x
      LdPlain<false> a{dgi + (long)d * 3 * H, 1, 6L * H, 0};
      LdPlain<false> b{x, 1, ldx, 0};
      int sk = pick_splitk(3 * H, In, M, 1);
      launch_gemm(a, b, make_epi(dw_ih[d], (long)In, nullptr, 0, 0.f, sk > 1 ? 2 : 1), 3 * H, In, M, 1, sk, stream);
      launch_colsum(dgi + (long)d * 3 * H, 6L * H, db_ih[d], M, 3 * H, stream);
    }
    {  // dW_hh[d][3H, H] += dgh_d^T @ h_prev_d  (h_prev = out shifted by one step, zero at the boundary)
      LdPlain<false> a{dgh + (long)d * 3 * H, 1, 6L * H, 0};
      LdT<LdConv<ORDER_KKC>> b{LdConv<ORDER_KKC>{out + (long)d * H, T, 1, H, T, 1, 1, 1, 1, 1, 1, 1, +1,
                                                 d == 0 ? -1 : +1, 0, 2L * H}};
      int sk = pick_splitk(3 * H, H, M, 1);
      launch_gemm(a, b, make_epi(dw_hh[d], (long)H, nullptr, 0, 0.f, sk > 1 ? 2 : 1), 3 * H, H, M, 1, sk, stream);
      launch_colsum(dgh + (long)d * 3 * H, 6L * H, db_hh[d], M, 3 * H, stream);
    }
  }
  if (dx && do_dx) {
    for (int d = 0; d < 2; ++d) {
      LdPlain<true> a{dgi + (long)d * 3 * H, 6L * H, 1, 0};
      LdPlain<false> b{w_ih[d], 1, (long)In, 0};
      launch_gemm(a, b, make_epi(dx, lddx, nullptr, 0, 0.f, d == 0 ? 0 : 1), M, In, 3 * H, 1, 1, stream);
    }
  }
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
