// Device-side long-form synthesis pipeline and evaluation metrics, sm_100a.
// All HBM-bound elementwise / small-reduction work: one thread per output element, coalesced along the pose dim.
//
//  s2ag_longform_blend     processor_v2.py:1282-1290 (seed hand-off: last n_pre frames of a chunk seed the next) and
//                          :1303-1327 (linear blend of the n_pre overlapping frames), for a whole batch of clips that
//                          advance in lock-step; clips whose own chunk count is exhausted are skipped.
//  s2ag_fade_out           :1334-1391 zero tail + weighted quadratic least-squares re-fit of the last 2*n_pre frames.
//  s2ag_dir_vec_to_pose    utils/ted_db_utils.py:81-102 (+ the mean direction vector, processor_v2.py:1413-1416).
//  s2ag_pose_metrics       processor_v2.py:738-774 `push_samples`: L1, joint MAE (frames >= n_pre), acceleration diff.
#include "common.cuh"

namespace {

#ifdef S2AG_EMU
__device__ inline float f_mul(float a, float b) { volatile float r = a * b; return r; }
__device__ inline float f_div(float a, float b) { volatile float r = a / b; return r; }
__device__ inline float f_add(float a, float b) { volatile float r = a + b; return r; }
#else
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
#endif

// utils/ted_db_utils.py:14 (parent, child, bone length); children appear after their parents
__device__ const int kParent[9] = {0, 1, 2, 1, 4, 5, 1, 7, 8};
__device__ const int kChild[9] = {1, 2, 3, 4, 5, 6, 7, 8, 9};
__device__ const float kBone[9] = {0.26f, 0.18f, 0.14f, 0.22f, 0.36f, 0.33f, 0.22f, 0.36f, 0.33f};

__global__ void longform_blend_kernel(const float* __restrict__ out, float* __restrict__ result, long ld_result,
                                      float* __restrict__ pre_next, const int* __restrict__ n_chunks, int chunk,
                                      int B, int T, int P, int n_pre) {
  const int stride = T - n_pre;
  const long total = (long)B * T * P;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int p = (int)(e % P);
    const int t = (int)((e / P) % T);
    const int b = (int)(e / ((long)P * T));
    if (n_chunks && chunk >= n_chunks[b]) continue;
    const float v = out[e];
    float* dst = result + (long)b * ld_result + ((long)chunk * stride + t) * P + p;
    if (chunk > 0 && t < n_pre) {
      // out[j] = prev[j] * (n - j) / (n + 1) + next[j] * (j + 1) / (n + 1), evaluated in fp32 in this order (numpy)
      const float n1 = (float)(n_pre + 1);
      *dst = f_add(f_div(f_mul(*dst, (float)(n_pre - t)), n1), f_div(f_mul(v, (float)(t + 1)), n1));
    } else {
      *dst = v;
    }
    if (pre_next && t >= T - n_pre) {   // seed of the next chunk: the RAW last n_pre frames (:1282-1290)
      float* q = pre_next + ((long)b * T + (t - (T - n_pre))) * (P + 1);
      q[p] = v;
      if (p == 0) q[P] = 1.f;
    }
  }
}

// one thread per (clip, pose dim): weighted quadratic fit over m = 2*n_smooth points, normal equations in fp64
__global__ void fade_out_kernel(float* __restrict__ seq, long ld_seq, const int* __restrict__ len,
                                const int* __restrict__ start_frame, int* __restrict__ len_out, int B, int P,
                                int n_smooth, int Lmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * P) return;
  const int b = i / P, p = i % P;
  const int s0 = start_frame[b], m = 2 * n_smooth, e0 = s0 + m;
  int L = len[b];
  float* y = seq + (long)b * ld_seq + p;
  if (e0 > Lmax || s0 < 0) S2AG_DEVICE_TRAP();
  if (L < e0) { for (int t = L; t < e0; ++t) y[(long)t * P] = 0.f; L = e0; }     // np.pad(..., 'constant')
  for (int t = e0 - n_smooth; t < L; ++t) y[(long)t * P] = 0.f;                   // fade to the mean pose
  double S[5] = {0, 0, 0, 0, 0}, R[3] = {0, 0, 0};
  for (int k = 0; k < m; ++k) {
    const double w = (k == 0 || k == m - 1) ? 5.0 : 1.0, w2 = w * w, x = (double)k;
    const double yy = (double)y[(long)(s0 + k) * P];
    double xp = 1.0;
    for (int q = 0; q < 5; ++q) { S[q] += w2 * xp; if (q < 3) R[q] += w2 * xp * yy; xp *= x; }
  }
  // solve [[S0 S1 S2],[S1 S2 S3],[S2 S3 S4]] c = R  (Cramer)
  const double a = S[0], bq = S[1], c = S[2], d = S[3], e = S[4];
  const double det = a * (c * e - d * d) - bq * (bq * e - d * c) + c * (bq * d - c * c);
  const double c0 = (R[0] * (c * e - d * d) - bq * (R[1] * e - d * R[2]) + c * (R[1] * d - c * R[2])) / det;
  const double c1 = (a * (R[1] * e - d * R[2]) - R[0] * (bq * e - d * c) + c * (bq * R[2] - R[1] * c)) / det;
  const double c2 = (a * (c * R[2] - R[1] * d) - bq * (bq * R[2] - R[1] * c) + R[0] * (bq * d - c * c)) / det;
  for (int k = 0; k < m; ++k) y[(long)(s0 + k) * P] = (float)(c0 + c1 * k + c2 * (double)k * k);
  if (p == 0 && len_out) len_out[b] = L;
}

__device__ __forceinline__ void joints_of(const float* __restrict__ v, const float* __restrict__ mean, double* J) {
  J[0] = J[1] = J[2] = 0.0;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = mean ? f_add(v[j * 3 + c], mean[j * 3 + c]) : v[j * 3 + c];
      J[kChild[j] * 3 + c] = J[kParent[j] * 3 + c] + (double)f_mul(kBone[j], d);
    }
  }
}

__global__ void dir_vec_to_pose_kernel(const float* __restrict__ vec, const float* __restrict__ mean,
                                       float* __restrict__ pose, long N) {
  for (long n = blockIdx.x * (long)blockDim.x + threadIdx.x; n < N; n += (long)gridDim.x * blockDim.x) {
    double J[30];
    joints_of(vec + n * 27, mean, J);
    for (int q = 0; q < 30; ++q) pose[n * 30 + q] = (float)J[q];
  }
}

// acc[0] = sum |out - tgt|, acc[1] = sum |J_out - J_tgt| over frames >= n_pre, acc[2] = sum |acc_tgt - acc_out|
__global__ void pose_metrics_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                    const float* __restrict__ mean, double* __restrict__ acc, int B, int T, int n_pre) {
  double l1 = 0.0, mae = 0.0, ac = 0.0;
  for (long n = blockIdx.x * (long)blockDim.x + threadIdx.x; n < (long)B * T; n += (long)gridDim.x * blockDim.x) {
    const int t = (int)(n % T);
    const float* o = out + n * 27;
    const float* g = tgt + n * 27;
    for (int q = 0; q < 27; ++q) l1 += fabs((double)o[q] - (double)g[q]);
    double Jo[30], Jt[30];
    joints_of(o, mean, Jo);
    joints_of(g, mean, Jt);
    if (t >= n_pre) for (int q = 0; q < 30; ++q) mae += fabs(Jo[q] - Jt[q]);
    if (t + 2 < T) {   // second difference over frames t, t+1, t+2
      double Jo1[30], Jt1[30], Jo2[30], Jt2[30];
      joints_of(o + 27, mean, Jo1); joints_of(g + 27, mean, Jt1);
      joints_of(o + 54, mean, Jo2); joints_of(g + 54, mean, Jt2);
      for (int q = 0; q < 30; ++q)
        ac += fabs((Jt2[q] - 2.0 * Jt1[q] + Jt[q]) - (Jo2[q] - 2.0 * Jo1[q] + Jo[q]));
    }
  }
  l1 = s2ag_warp_sum_d(l1); mae = s2ag_warp_sum_d(mae); ac = s2ag_warp_sum_d(ac);
  if ((threadIdx.x & 31) == 0) { atomicAdd(acc + 0, l1); atomicAdd(acc + 1, mae); atomicAdd(acc + 2, ac); }
}

__global__ void pose_metrics_final_kernel(const double* __restrict__ acc, float* __restrict__ dst, double n0, double n1,
                                          double n2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    dst[0] = (float)(acc[0] / n0);
    dst[1] = (float)(acc[1] / n1);
    dst[2] = (float)(acc[2] / n2);
  }
}

}  // namespace

extern "C" int s2ag_longform_blend(const float* out, float* result, long ld_result, float* pre_next,
                                   const int* n_chunks, int chunk, int B, int T, int P, int n_pre, void* stream) {
  S2AG_CHECK_ARG(out && result && B >= 0 && T > 0 && P > 0 && n_pre >= 0 && n_pre < T && chunk >= 0);
  S2AG_CHECK_ARG(ld_result >= ((long)chunk * (T - n_pre) + T) * P);
  if (B == 0) return S2AG_OK;
  if (pre_next) {
    cudaError_t e = cudaMemsetAsync(pre_next, 0, (size_t)B * T * (P + 1) * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) { s2ag_set_error("memset failed"); return S2AG_ERR_LAUNCH; }
  }
  long total = (long)B * T * P;
  long blocks = (total + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
  auto k = &longform_blend_kernel;
  S2AG_LAUNCH(k, (int)blocks, 256, 0, stream, out, result, ld_result, pre_next, n_chunks, chunk, B, T, P, n_pre);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_fade_out(float* seq, long ld_seq, const int* len, const int* start_frame, int* len_out, int B, int P,
                             int n_smooth, int Lmax, void* stream) {
  S2AG_CHECK_ARG(seq && len && start_frame && B >= 0 && P > 0 && n_smooth > 0 && Lmax > 0 && ld_seq >= (long)Lmax * P);
  if (B == 0) return S2AG_OK;
  auto k = &fade_out_kernel;
  S2AG_LAUNCH(k, s2ag_cdiv((long)B * P, 128), 128, 0, stream, seq, ld_seq, len, start_frame, len_out, B, P, n_smooth, Lmax);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_dir_vec_to_pose(const float* vec, const float* mean, float* pose, long N, void* stream) {
  S2AG_CHECK_ARG(vec && pose && N >= 0);
  if (N == 0) return S2AG_OK;
  long blocks = (N + 127) / 128; if (blocks > 148 * 16) blocks = 148 * 16;
  auto k = &dir_vec_to_pose_kernel;
  S2AG_LAUNCH(k, (int)blocks, 128, 0, stream, vec, mean, pose, N);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_pose_metrics(const float* out, const float* tgt, const float* mean, double* acc_ws, float* dst,
                                 int B, int T, int n_pre, void* stream) {
  S2AG_CHECK_ARG(out && tgt && acc_ws && dst && B > 0 && T > 2 && n_pre >= 0 && n_pre < T);
  cudaError_t e = cudaMemsetAsync(acc_ws, 0, 3 * sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) { s2ag_set_error("memset failed"); return S2AG_ERR_LAUNCH; }
  long blocks = ((long)B * T + 127) / 128; if (blocks > 148 * 8) blocks = 148 * 8;
  auto k = &pose_metrics_kernel;
  S2AG_LAUNCH(k, (int)blocks, 128, 0, stream, out, tgt, mean, acc_ws, B, T, n_pre);
  S2AG_CHECK_LAUNCH();
  auto k2 = &pose_metrics_final_kernel;
  S2AG_LAUNCH(k2, 1, 32, 0, stream, (const double*)acc_ws, dst, (double)B * T * 27, (double)B * (T - n_pre) * 30,
              (double)B * (T - 2) * 30);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
