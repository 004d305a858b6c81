// BatchNorm (train / eval) over channels-last [M, C] activations, fused with the activation that
// follows it, an optional residual add, and an optional output-column permutation.
// Replaces nn.BatchNorm1d/2d (+ LeakyReLU/ReLU, + the AffEncoder view/permute regrouping):
//   reference net/multimodal_context_net_v2.py:19-25,40-46,123,131,144,149,153-175; net/utils/tgcn.py:179,188,204,217.
// HBM-bound: forward = 2 reads + 1 write of the activation, backward = 3 reads (+1: y) + 1 write.
// Statistics are accumulated as shifted sums (shift = first row) in fp32 per thread and fp64 across
// threads/blocks, so var = E[(x-s)^2] - E[x-s]^2 has no catastrophic cancellation.
#include "s2ag.h"
#include "common.cuh"

// A/B switch (s2ag_debug_bn_flags): bit 0 = scalar kernels everywhere (no float4 variants)
static int g_bn_flags = 0;
extern "C" int s2ag_debug_bn_flags(int flags) { int o = g_bn_flags; g_bn_flags = flags; return o; }

namespace {

struct BnGeom { int cb; int rif; };  // columns per block (8/16/32), rows in flight (256/cb)
static inline BnGeom bn_geom(int C) {
  BnGeom g; g.cb = C <= 8 ? 8 : (C <= 16 ? 16 : 32); g.rif = 256 / g.cb; return g;
}
static inline int bn_rows_per_block(int M, int colblocks) {
  // aim for >= 4 waves of 148 SMs but keep >= 64 rows per block
  long want_blocks = 148 * 4 / (colblocks > 0 ? colblocks : 1) + 1;
  long rpb = (M + want_blocks - 1) / want_blocks;
  if (rpb < 64) rpb = 64;
  if (rpb > 4096) rpb = 4096;
  return (int)rpb;
}

// ws[c] += sum_m (x[m,c]-x[0,c]) ; ws[C+c] += sum_m (x[m,c]-x[0,c])^2
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, long ldx, int M, int C,
                                                       double* __restrict__ ws, int cb, int rpb) {
  __shared__ double r1[256];
  __shared__ double r2[256];
  // blockIdx.z = statistics group (M rows each): independent batches that share the weights (see s2ag_bn_fwd)
  x += (long)blockIdx.z * M * ldx; ws += (long)blockIdx.z * 2 * C;
  const int tx = threadIdx.x % cb, ty = threadIdx.x / cb, rif = 256 / cb;
  const int c = blockIdx.x * cb + tx;
  const int mbeg = blockIdx.y * rpb;
  int mend = mbeg + rpb; if (mend > M) mend = M;
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    const float shift = __ldg(x + c);
    for (int m = mbeg + ty; m < mend; m += rif) {
      float v = __ldg(x + (long)m * ldx + c) - shift;
      s1 += v; s2 = fmaf(v, v, s2);
    }
  }
  r1[threadIdx.x] = (double)s1; r2[threadIdx.x] = (double)s2;
  __syncthreads();
  if (ty == 0 && c < C) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < rif; ++i) { a += r1[i * cb + tx]; b += r2[i * cb + tx]; }
    atomicAdd(ws + c, a); atomicAdd(ws + C + c, b);
  }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(
    const float* __restrict__ x, long ldx, int M, int C, const float* __restrict__ gamma,
    const float* __restrict__ beta, const int32_t* __restrict__ pmap, float* __restrict__ rmean,
    float* __restrict__ rvar, int training, float momentum, float eps, const float* __restrict__ add, long ldadd,
    float* __restrict__ y, long ldy, const int32_t* __restrict__ cmap, int act, float slope,
    float* __restrict__ save_mean, float* __restrict__ save_invstd, const double* __restrict__ ws, int cb, int rpb) {
  const int tx = threadIdx.x % cb, ty = threadIdx.x / cb, rif = 256 / cb;
  const int c = blockIdx.x * cb + tx;
  if (c >= C) return;  // no barriers below
  const int pi = pmap ? pmap[c] : c;
  const int oc = cmap ? cmap[c] : c;
  const int grp = blockIdx.z, groups = gridDim.z;
  const float* x0 = x;   // group 0 (running statistics are updated for every group, in order, by group 0's block)
  x += (long)grp * M * ldx; y += (long)grp * M * ldy;
  if (add) add += (long)grp * M * ldadd;
  if (save_mean) save_mean += (long)grp * C;
  if (save_invstd) save_invstd += (long)grp * C;
  float mean, invstd;
  if (training) {
    const double* wg = ws + (long)grp * 2 * C;
    const double shift = (double)__ldg(x + c);
    const double e1 = wg[c] / (double)M, e2 = wg[C + c] / (double)M;
    double var = e2 - e1 * e1; if (var < 0.0) var = 0.0;
    mean = (float)(shift + e1);
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (blockIdx.y == 0 && ty == 0) {
      if (save_mean) save_mean[c] = mean;
      if (save_invstd) save_invstd[c] = invstd;
      if (rmean && grp == 0) {
        // momentum updates of the groups in call order (the reference runs the module once per group)
        float rm = rmean[pi], rv = rvar[pi];
        for (int q = 0; q < groups; ++q) {
          const double sh = (double)__ldg(x0 + (long)q * M * ldx + c);
          const double q1 = ws[(long)q * 2 * C + c] / (double)M, q2 = ws[(long)q * 2 * C + C + c] / (double)M;
          double vq = q2 - q1 * q1; if (vq < 0.0) vq = 0.0;
          const float unbiased = (float)(M > 1 ? vq * (double)M / (double)(M - 1) : vq);
          rm = (1.f - momentum) * rm + momentum * (float)(sh + q1);
          rv = (1.f - momentum) * rv + momentum * unbiased;
        }
        rmean[pi] = rm; rvar[pi] = rv;
      }
    }
  } else {
    mean = rmean[pi];
    invstd = 1.f / sqrtf(rvar[pi] + eps);
    if (blockIdx.y == 0 && ty == 0) {
      if (save_mean) save_mean[c] = mean;
      if (save_invstd) save_invstd[c] = invstd;
    }
  }
  const float g = gamma ? gamma[pi] : 1.f, b = beta ? beta[pi] : 0.f;
  const float scale = g * invstd, bias = b - mean * scale;
  const int mbeg = blockIdx.y * rpb;
  int mend = mbeg + rpb; if (mend > M) mend = M;
  for (int m = mbeg + ty; m < mend; m += rif) {
    float v = fmaf(__ldg(x + (long)m * ldx + c), scale, bias);
    if (add) v += __ldg(add + (long)m * ldadd + oc);
    y[(long)m * ldy + oc] = s2ag_act(v, act, slope);
  }
}

// ws[c] += sum dpre ; ws[C+c] += sum dpre * xhat ;  dadd = dpre
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(
    const float* __restrict__ dy, long lddy, const float* __restrict__ y, long ldy, const int32_t* __restrict__ cmap,
    const float* __restrict__ x, long ldx, int M, int C, const float* __restrict__ save_mean,
    const float* __restrict__ save_invstd, int act, float slope, float* __restrict__ dadd, long lddadd,
    double* __restrict__ ws, int cb, int rpb) {
  __shared__ double r1[256];
  __shared__ double r2[256];
  {
    const long grp = blockIdx.z;
    dy += grp * M * lddy; x += grp * M * ldx; save_mean += grp * C; save_invstd += grp * C; ws += grp * 2 * C;
    if (y) y += grp * M * ldy;
    if (dadd) dadd += grp * M * lddadd;
  }
  const int tx = threadIdx.x % cb, ty = threadIdx.x / cb, rif = 256 / cb;
  const int c = blockIdx.x * cb + tx;
  const int mbeg = blockIdx.y * rpb;
  int mend = mbeg + rpb; if (mend > M) mend = M;
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    const int oc = cmap ? cmap[c] : c;
    const float mean = save_mean[c], invstd = save_invstd[c];
    for (int m = mbeg + ty; m < mend; m += rif) {
      float d = __ldg(dy + (long)m * lddy + oc);
      if (act != S2AG_ACT_NONE) d *= s2ag_act_grad_from_out(__ldg(y + (long)m * ldy + oc), act, slope);
      if (dadd) dadd[(long)m * lddadd + oc] = d;
      const float xh = (__ldg(x + (long)m * ldx + c) - mean) * invstd;
      s1 += d; s2 = fmaf(d, xh, s2);
    }
  }
  r1[threadIdx.x] = (double)s1; r2[threadIdx.x] = (double)s2;
  __syncthreads();
  if (ty == 0 && c < C) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < rif; ++i) { a += r1[i * cb + tx]; b += r2[i * cb + tx]; }
    atomicAdd(ws + c, a); atomicAdd(ws + C + c, b);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(
    const float* __restrict__ dy, long lddy, const float* __restrict__ y, long ldy, const int32_t* __restrict__ cmap,
    const float* __restrict__ x, long ldx, int M, int C, const float* __restrict__ gamma,
    const int32_t* __restrict__ pmap, const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
    int training, int act, float slope, float* __restrict__ dx, long lddx, float* __restrict__ dgamma,
    float* __restrict__ dbeta, const double* __restrict__ ws, int cb, int rpb) {
  const int tx = threadIdx.x % cb, ty = threadIdx.x / cb, rif = 256 / cb;
  const int c = blockIdx.x * cb + tx;
  if (c >= C) return;
  const int pi = pmap ? pmap[c] : c;
  const int oc = cmap ? cmap[c] : c;
  const long grp = blockIdx.z;
  if (blockIdx.y == 0 && ty == 0 && grp == 0) {   // parameter gradients: sum over the groups, one writer
    double sd = 0.0, sx = 0.0;
    for (int q = 0; q < (int)gridDim.z; ++q) { sd += ws[(long)q * 2 * C + c]; sx += ws[(long)q * 2 * C + C + c]; }
    // atomics: two backward chains of the same module may run on different streams (D(real) / D(fake))
    if (dgamma) atomicAdd(dgamma + pi, (float)sx);
    if (dbeta) atomicAdd(dbeta + pi, (float)sd);
  }
  dy += grp * M * lddy; x += grp * M * ldx; save_mean += grp * C; save_invstd += grp * C; ws += grp * 2 * C;
  if (y) y += grp * M * ldy;
  if (dx) dx += grp * M * lddx;
  const float mean = save_mean[c], invstd = save_invstd[c];
  const float g = gamma ? gamma[pi] : 1.f;
  const float sum_d = (float)ws[c], sum_dx = (float)ws[C + c];
  if (!dx) return;
  const float k1 = training ? sum_d / (float)M : 0.f, k2 = training ? sum_dx / (float)M : 0.f;
  const float gs = g * invstd;
  const int mbeg = blockIdx.y * rpb;
  int mend = mbeg + rpb; if (mend > M) mend = M;
  for (int m = mbeg + ty; m < mend; m += rif) {
    float d = __ldg(dy + (long)m * lddy + oc);
    if (act != S2AG_ACT_NONE) d *= s2ag_act_grad_from_out(__ldg(y + (long)m * ldy + oc), act, slope);
    const float xh = (__ldg(x + (long)m * ldx + c) - mean) * invstd;
    dx[(long)m * lddx + c] = gs * (d - k1 - xh * k2);
  }
}

// ---- float4 variants: 4 channels per thread (one 16-byte load per operand and row), used when there is no output-column
// permutation, C % 4 == 0 and every row pitch / base pointer is 16-byte aligned.  Same statistics convention, same
// workspace layout and the same (cb, rpb) grid logic with cb counted in float4 columns.
struct BnGeomV { int cbv; };
static inline BnGeomV bn_geom_v(int C) {
  int cv = C / 4; BnGeomV g; g.cbv = cv <= 1 ? 1 : (cv <= 2 ? 2 : (cv <= 4 ? 4 : 8)); return g;
}
static inline bool bn_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__device__ __forceinline__ void bn_reduce4(double (*r)[256], const float (&s)[8], int cbv, int tx, int ty, int c4, int C,
                                           double* ws) {
  // r[0..7][256]: 4 channels x 2 sums per thread
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i][threadIdx.x] = (double)s[i];
  __syncthreads();
  const int rif = 256 / cbv;
  if (ty < 8 && c4 * 4 < C) {   // 8 helper rows per column: one (component, sum) pair each
    const int comp = ty & 3, which = ty >> 2;
    double a = 0.0;
    for (int i = 0; i < rif; ++i) a += r[which * 4 + comp][i * cbv + tx];
    atomicAdd(ws + (which ? C : 0) + c4 * 4 + comp, a);
  }
}

__global__ void __launch_bounds__(256) bn_stats_v4_kernel(const float* __restrict__ x, long ldx, int M, int C,
                                                          double* __restrict__ ws, int cbv, int rpb) {
  __shared__ double r[8][256];
  x += (long)blockIdx.z * M * ldx; ws += (long)blockIdx.z * 2 * C;
  const int tx = threadIdx.x % cbv, ty = threadIdx.x / cbv, rif = 256 / cbv;
  const int c4 = blockIdx.x * cbv + tx;
  const int mbeg = blockIdx.y * rpb;
  int mend = mbeg + rpb; if (mend > M) mend = M;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c4 * 4 < C) {
    const float4 sh = __ldg(reinterpret_cast<const float4*>(x) + c4);
#pragma unroll 4
    for (int m = mbeg + ty; m < mend; m += rif) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long)m * ldx) + c4);
      const float a = v.x - sh.x, b = v.y - sh.y, c = v.z - sh.z, d = v.w - sh.w;
      s[0] += a; s[1] += b; s[2] += c; s[3] += d;
      s[4] = fmaf(a, a, s[4]); s[5] = fmaf(b, b, s[5]); s[6] = fmaf(c, c, s[6]); s[7] = fmaf(d, d, s[7]);
    }
  }
  bn_reduce4(r, s, cbv, tx, ty, c4, C, ws);
}

__global__ void __launch_bounds__(256) bn_apply_v4_kernel(
    const float* __restrict__ x, long ldx, int M, int C, const float* __restrict__ gamma,
    const float* __restrict__ beta, const int32_t* __restrict__ pmap, float* __restrict__ rmean,
    float* __restrict__ rvar, int training, float momentum, float eps, const float* __restrict__ add, long ldadd,
    float* __restrict__ y, long ldy, int act, float slope, float* __restrict__ save_mean,
    float* __restrict__ save_invstd, const double* __restrict__ ws, int cbv, int rpb) {
  const int tx = threadIdx.x % cbv, ty = threadIdx.x / cbv, rif = 256 / cbv;
  const int c4 = blockIdx.x * cbv + tx;
  if (c4 * 4 >= C) return;  // no barriers below
  const int grp = blockIdx.z, groups = gridDim.z;
  const float* x0 = x;
  x += (long)grp * M * ldx; y += (long)grp * M * ldy;
  if (add) add += (long)grp * M * ldadd;
  if (save_mean) save_mean += (long)grp * C;
  if (save_invstd) save_invstd += (long)grp * C;
  float scale[4], bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c4 * 4 + j;
    const int pi = pmap ? pmap[c] : c;
    float mean, invstd;
    if (training) {
      const double* wg = ws + (long)grp * 2 * C;
      const double shift = (double)__ldg(x + c);
      const double e1 = wg[c] / (double)M, e2 = wg[C + c] / (double)M;
      double var = e2 - e1 * e1; if (var < 0.0) var = 0.0;
      mean = (float)(shift + e1);
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.y == 0 && ty == 0 && rmean && grp == 0) {
        float rm = rmean[pi], rv = rvar[pi];
        for (int q = 0; q < groups; ++q) {
          const double sh = (double)__ldg(x0 + (long)q * M * ldx + c);
          const double q1 = ws[(long)q * 2 * C + c] / (double)M, q2 = ws[(long)q * 2 * C + C + c] / (double)M;
          double vq = q2 - q1 * q1; if (vq < 0.0) vq = 0.0;
          const float unbiased = (float)(M > 1 ? vq * (double)M / (double)(M - 1) : vq);
          rm = (1.f - momentum) * rm + momentum * (float)(sh + q1);
          rv = (1.f - momentum) * rv + momentum * unbiased;
        }
        rmean[pi] = rm; rvar[pi] = rv;
      }
    } else {
      mean = rmean[pi];
      invstd = 1.f / sqrtf(rvar[pi] + eps);
    }
    if (blockIdx.y == 0 && ty == 0) {
      if (save_mean) save_mean[c] = mean;
      if (save_invstd) save_invstd[c] = invstd;
    }
    const float g = gamma ? gamma[pi] : 1.f, b = beta ? beta[pi] : 0.f;
    scale[j] = g * invstd; bias[j] = b - mean * scale[j];
  }
  const int mbeg = blockIdx.y * rpb;
  int mend = mbeg + rpb; if (mend > M) mend = M;
#pragma unroll 4
  for (int m = mbeg + ty; m < mend; m += rif) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long)m * ldx) + c4);
    float4 o;
    o.x = fmaf(v.x, scale[0], bias[0]); o.y = fmaf(v.y, scale[1], bias[1]);
    o.z = fmaf(v.z, scale[2], bias[2]); o.w = fmaf(v.w, scale[3], bias[3]);
    if (add) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(add + (long)m * ldadd) + c4);
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    o.x = s2ag_act(o.x, act, slope); o.y = s2ag_act(o.y, act, slope);
    o.z = s2ag_act(o.z, act, slope); o.w = s2ag_act(o.w, act, slope);
    reinterpret_cast<float4*>(y + (long)m * ldy)[c4] = o;
  }
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_v4_kernel(
    const float* __restrict__ dy, long lddy, const float* __restrict__ y, long ldy, const float* __restrict__ x, long ldx,
    int M, int C, const float* __restrict__ save_mean, const float* __restrict__ save_invstd, int act, float slope,
    float* __restrict__ dadd, long lddadd, double* __restrict__ ws, int cbv, int rpb) {
  __shared__ double r[8][256];
  {
    const long grp = blockIdx.z;
    dy += grp * M * lddy; x += grp * M * ldx; save_mean += grp * C; save_invstd += grp * C; ws += grp * 2 * C;
    if (y) y += grp * M * ldy;
    if (dadd) dadd += grp * M * lddadd;
  }
  const int tx = threadIdx.x % cbv, ty = threadIdx.x / cbv, rif = 256 / cbv;
  const int c4 = blockIdx.x * cbv + tx;
  const int mbeg = blockIdx.y * rpb;
  int mend = mbeg + rpb; if (mend > M) mend = M;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c4 * 4 < C) {
    const float4 mean = reinterpret_cast<const float4*>(save_mean)[c4], inv = reinterpret_cast<const float4*>(save_invstd)[c4];
#pragma unroll 2
    for (int m = mbeg + ty; m < mend; m += rif) {
      float4 d = __ldg(reinterpret_cast<const float4*>(dy + (long)m * lddy) + c4);
      if (act != S2AG_ACT_NONE) {
        const float4 o = __ldg(reinterpret_cast<const float4*>(y + (long)m * ldy) + c4);
        d.x *= s2ag_act_grad_from_out(o.x, act, slope); d.y *= s2ag_act_grad_from_out(o.y, act, slope);
        d.z *= s2ag_act_grad_from_out(o.z, act, slope); d.w *= s2ag_act_grad_from_out(o.w, act, slope);
      }
      if (dadd) reinterpret_cast<float4*>(dadd + (long)m * lddadd)[c4] = d;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long)m * ldx) + c4);
      s[0] += d.x; s[1] += d.y; s[2] += d.z; s[3] += d.w;
      s[4] = fmaf(d.x, (v.x - mean.x) * inv.x, s[4]); s[5] = fmaf(d.y, (v.y - mean.y) * inv.y, s[5]);
      s[6] = fmaf(d.z, (v.z - mean.z) * inv.z, s[6]); s[7] = fmaf(d.w, (v.w - mean.w) * inv.w, s[7]);
    }
  }
  bn_reduce4(r, s, cbv, tx, ty, c4, C, ws);
}

__global__ void __launch_bounds__(256) bn_bwd_apply_v4_kernel(
    const float* __restrict__ dy, long lddy, const float* __restrict__ y, long ldy, const float* __restrict__ x, long ldx,
    int M, int C, const float* __restrict__ gamma, const int32_t* __restrict__ pmap,
    const float* __restrict__ save_mean, const float* __restrict__ save_invstd, int training, int act, float slope,
    float* __restrict__ dx, long lddx, float* __restrict__ dgamma, float* __restrict__ dbeta,
    const double* __restrict__ ws, int cbv, int rpb) {
  const int tx = threadIdx.x % cbv, ty = threadIdx.x / cbv, rif = 256 / cbv;
  const int c4 = blockIdx.x * cbv + tx;
  if (c4 * 4 >= C) return;
  const long grp = blockIdx.z;
  if (blockIdx.y == 0 && ty == 0 && grp == 0) {   // parameter gradients: sum over the groups, one writer
    for (int j = 0; j < 4; ++j) {
      const int c = c4 * 4 + j;
      const int pi = pmap ? pmap[c] : c;
      double sd = 0.0, sx = 0.0;
      for (int q = 0; q < (int)gridDim.z; ++q) { sd += ws[(long)q * 2 * C + c]; sx += ws[(long)q * 2 * C + C + c]; }
      if (dgamma) atomicAdd(dgamma + pi, (float)sx);
      if (dbeta) atomicAdd(dbeta + pi, (float)sd);
    }
  }
  if (!dx) return;
  dy += grp * M * lddy; x += grp * M * ldx; save_mean += grp * C; save_invstd += grp * C; ws += grp * 2 * C;
  if (y) y += grp * M * ldy;
  dx += grp * M * lddx;
  float mean[4], inv[4], gs[4], k1[4], k2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c4 * 4 + j;
    const int pi = pmap ? pmap[c] : c;
    mean[j] = save_mean[c]; inv[j] = save_invstd[c];
    gs[j] = (gamma ? gamma[pi] : 1.f) * inv[j];
    k1[j] = training ? (float)ws[c] / (float)M : 0.f;
    k2[j] = training ? (float)ws[C + c] / (float)M : 0.f;
  }
  const int mbeg = blockIdx.y * rpb;
  int mend = mbeg + rpb; if (mend > M) mend = M;
#pragma unroll 2
  for (int m = mbeg + ty; m < mend; m += rif) {
    float4 d = __ldg(reinterpret_cast<const float4*>(dy + (long)m * lddy) + c4);
    if (act != S2AG_ACT_NONE) {
      const float4 o = __ldg(reinterpret_cast<const float4*>(y + (long)m * ldy) + c4);
      d.x *= s2ag_act_grad_from_out(o.x, act, slope); d.y *= s2ag_act_grad_from_out(o.y, act, slope);
      d.z *= s2ag_act_grad_from_out(o.z, act, slope); d.w *= s2ag_act_grad_from_out(o.w, act, slope);
    }
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long)m * ldx) + c4);
    float4 o;
    o.x = gs[0] * (d.x - k1[0] - (v.x - mean[0]) * inv[0] * k2[0]);
    o.y = gs[1] * (d.y - k1[1] - (v.y - mean[1]) * inv[1] * k2[1]);
    o.z = gs[2] * (d.z - k1[2] - (v.z - mean[2]) * inv[2] * k2[2]);
    o.w = gs[3] * (d.w - k1[3] - (v.w - mean[3]) * inv[3] * k2[3]);
    reinterpret_cast<float4*>(dx + (long)m * lddx)[c4] = o;
  }
}

}  // namespace

extern "C" int s2ag_bn_fwd(const float* x, long ldx, int M, int C, const float* gamma, const float* beta,
                           const int32_t* param_map, float* running_mean, float* running_var,
                           int training, float momentum, float eps,
                           const float* add, long ldadd, float* y, long ldy, const int32_t* col_map,
                           int act, float slope, float* save_mean, float* save_invstd, double* ws, int groups,
                           void* stream) {
  S2AG_CHECK_ARG(x && y && M > 0 && C > 0 && ldx >= C && ldy >= C && groups >= 1 && M % groups == 0);
  S2AG_CHECK_ARG(training ? (ws != nullptr) : (running_mean && running_var));
  M /= groups;   // rows per statistics group
  if (!col_map && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && (!add || ldadd % 4 == 0) && bn_al16(x) && bn_al16(y) &&
      bn_al16(add) && !(g_bn_flags & 1)) {
    BnGeomV gv = bn_geom_v(C);
    int colblocks = s2ag_cdiv(C / 4, gv.cbv);
    int rpb = bn_rows_per_block(M, colblocks * groups);
    dim3 grid(colblocks, s2ag_cdiv(M, rpb), groups);
    if (training) {
      cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C * groups, (cudaStream_t)stream);
      auto k1 = &bn_stats_v4_kernel;
      S2AG_LAUNCH(k1, grid, 256, 0, stream, x, ldx, M, C, ws, gv.cbv, rpb);
    }
    auto k2 = &bn_apply_v4_kernel;
    S2AG_LAUNCH(k2, grid, 256, 0, stream, x, ldx, M, C, gamma, beta, param_map, running_mean, running_var, training,
                momentum, eps, add, ldadd, y, ldy, act, slope, save_mean, save_invstd, (const double*)ws, gv.cbv, rpb);
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
  BnGeom g = bn_geom(C);
  int colblocks = s2ag_cdiv(C, g.cb);
  int rpb = bn_rows_per_block(M, colblocks * groups);
  dim3 grid(colblocks, s2ag_cdiv(M, rpb), groups);
  if (training) {
    cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C * groups, (cudaStream_t)stream);
    auto k1 = &bn_stats_kernel;
    S2AG_LAUNCH(k1, grid, 256, 0, stream, x, ldx, M, C, ws, g.cb, rpb);
  }
  auto k2 = &bn_apply_kernel;
  S2AG_LAUNCH(k2, grid, 256, 0, stream, x, ldx, M, C, gamma, beta, param_map, running_mean, running_var, training,
              momentum, eps, add, ldadd, y, ldy, col_map, act, slope, save_mean, save_invstd, (const double*)ws, g.cb,
              rpb);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_bn_bwd(const float* dy, long lddy, const float* y, long ldy, const int32_t* col_map,
                           const float* x, long ldx, int M, int C, const float* gamma, const int32_t* param_map,
                           const float* save_mean, const float* save_invstd, int training, int act, float slope,
                           float* dx, long lddx, float* dgamma, float* dbeta, float* dadd, long lddadd,
                           double* ws, int groups, void* stream) {
  S2AG_CHECK_ARG(dy && x && M > 0 && C > 0 && save_mean && save_invstd && ws && groups >= 1 && M % groups == 0);
  S2AG_CHECK_ARG(act == S2AG_ACT_NONE || y != nullptr);
  M /= groups;
  if (!col_map && C % 4 == 0 && lddy % 4 == 0 && ldx % 4 == 0 && (!y || ldy % 4 == 0) && (!dx || lddx % 4 == 0) &&
      (!dadd || lddadd % 4 == 0) && bn_al16(dy) && bn_al16(y) && bn_al16(x) && bn_al16(dx) && bn_al16(dadd) &&
      bn_al16(save_mean) && bn_al16(save_invstd) && !(g_bn_flags & 1)) {
    BnGeomV gv = bn_geom_v(C);
    int colblocks = s2ag_cdiv(C / 4, gv.cbv);
    int rpb = bn_rows_per_block(M, colblocks * groups);
    dim3 grid(colblocks, s2ag_cdiv(M, rpb), groups);
    cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C * groups, (cudaStream_t)stream);
    auto k1 = &bn_bwd_reduce_v4_kernel;
    S2AG_LAUNCH(k1, grid, 256, 0, stream, dy, lddy, y, ldy, x, ldx, M, C, save_mean, save_invstd, act, slope, dadd,
                lddadd, ws, gv.cbv, rpb);
    auto k2 = &bn_bwd_apply_v4_kernel;
    S2AG_LAUNCH(k2, grid, 256, 0, stream, dy, lddy, y, ldy, x, ldx, M, C, gamma, param_map, save_mean, save_invstd,
                training, act, slope, dx, lddx, dgamma, dbeta, (const double*)ws, gv.cbv, rpb);
    S2AG_CHECK_LAUNCH();
    return S2AG_OK;
  }
  BnGeom g = bn_geom(C);
  int colblocks = s2ag_cdiv(C, g.cb);
  int rpb = bn_rows_per_block(M, colblocks * groups);
  dim3 grid(colblocks, s2ag_cdiv(M, rpb), groups);
  cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C * groups, (cudaStream_t)stream);
  auto k1 = &bn_bwd_reduce_kernel;
  S2AG_LAUNCH(k1, grid, 256, 0, stream, dy, lddy, y, ldy, col_map, x, ldx, M, C, save_mean, save_invstd, act, slope,
              dadd, lddadd, ws, g.cb, rpb);
  auto k2 = &bn_bwd_apply_kernel;
  S2AG_LAUNCH(k2, grid, 256, 0, stream, dy, lddy, y, ldy, col_map, x, ldx, M, C, gamma, param_map, save_mean,
              save_invstd, training, act, slope, dx, lddx, dgamma, dbeta, (const double*)ws, g.cb, rpb);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
