// BatchNorm (train / eval) over channels-last [M, C] activations, fused with the activation that
// follows it, an optional residual add, and an optional output-column permutation.
// Replaces nn.BatchNorm1d/2d (+ LeakyReLU/ReLU, + the AffEncoder view/permute regrouping):
//   reference net/multimodal_context_net_v2.py:19-25,40-46,123,131,144,149,153-175; net/utils/tgcn.py:179,188,204,217.
// HBM-bound: forward = 2 reads + 1 write of the activation, backward = 3 reads (+1: y) + 1 write.
// Statistics are accumulated as shifted sums (shift = first row) in fp32 per thread and fp64 across
// threads/blocks, so var = E[(x-s)^2] - E[x-s]^2 has no catastrophic cancellation.
#include "s2ag.h"
#include "common.cuh"

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
    if (dgamma) dgamma[pi] += (float)sx;
    if (dbeta) dbeta[pi] += (float)sd;
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
