// Fréchet gesture distance on the device.
//
// Reference: net/embedding_space_evaluator.py:45-61 (push_samples keeps every latent feature of the real and the
// generated clips in host lists), :73-101 (get_scores: np.mean / np.cov over the stacked features, the Fréchet distance,
// and the mean per-sample L1 distance between paired features) and :104-152 (calculate_frechet_distance:
// ||mu1 - mu2||^2 + Tr(S1) + Tr(S2) - 2 Tr(sqrtm(S1 S2)) with scipy.linalg.sqrtm on the host).
//
// Here nothing goes back to the host per batch: `s2ag_fgd_accumulate` folds a batch of feature pairs into a small fp64
// moment buffer (count, paired L1 sum, per-dimension sums, second-moment matrices), and `s2ag_fgd_scores` turns the buffer
// into (frechet_dist, feat_dist) with one warp.  Tr(sqrtm(S1 S2)) is evaluated without a non-symmetric square root:
// the eigenvalues of S1 S2 are those of the symmetric PSD matrix S1^(1/2) S2 S1^(1/2), so two cyclic Jacobi
// eigen-decompositions (fp64, D <= 32: lane k owns row / column k) give Tr = sum_k sqrt(max(lambda_k, 0)).  The
// reference's fall-backs (eps on the diagonal when sqrtm returns non-finite values, ValueError on an imaginary
// component -> 1e10) cannot trigger on this route; a non-finite input moment still propagates to the result.
#include "common.cuh"

namespace {

constexpr int kMaxD = 32;
constexpr int kP = kMaxD + 1;   // padded pitch of the shared matrices
constexpr int kRows = 32;       // feature rows per staged tile

// moment buffer (doubles): [0] n, [1] sum_i |real_i - gen_i|_1, [2, 2+D) sum gen, [2+D, 2+2D) sum real,
// then D*D second moments of gen, D*D of real
__device__ __forceinline__ long acc_sum_g(int) { return 2; }
__device__ __forceinline__ long acc_sum_r(int D) { return 2 + D; }
__device__ __forceinline__ long acc_mom_g(int D) { return 2 + 2 * D; }
__device__ __forceinline__ long acc_mom_r(int D) { return 2 + 2 * D + (long)D * D; }

__global__ void __launch_bounds__(256) fgd_accumulate_kernel(const float* __restrict__ gen, long ldg,
                                                              const float* __restrict__ real, long ldr, int N, int D,
                                                              double* __restrict__ acc) {
  __shared__ float sg[kRows][kP], sr[kRows][kP];
  const int tid = threadIdx.x;
  double mg[4] = {0, 0, 0, 0}, mr[4] = {0, 0, 0, 0}, colsum = 0, l1 = 0;
  for (long r0 = (long)blockIdx.x * kRows; r0 < N; r0 += (long)gridDim.x * kRows) {
    int rows = (int)((N - r0) < kRows ? (N - r0) : kRows);
    for (int e = tid; e < kRows * D; e += 256) {
      int r = e / D, c = e - r * D;
      bool ok = r < rows;
      sg[r][c] = ok ? gen[(r0 + r) * ldg + c] : 0.f;
      sr[r][c] = ok ? real[(r0 + r) * ldr + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + q * 256;
      if (e < D * D) {
        int i = e / D, j = e - i * D;
        double a = 0, b = 0;
        for (int r = 0; r < kRows; ++r) {
          a += (double)sg[r][i] * (double)sg[r][j];
          b += (double)sr[r][i] * (double)sr[r][j];
        }
        mg[q] += a; mr[q] += b;
      }
    }
    if (tid < D) {
      for (int r = 0; r < kRows; ++r) colsum += (double)sg[r][tid];
    } else if (tid >= 32 && tid < 32 + D) {
      for (int r = 0; r < kRows; ++r) colsum += (double)sr[r][tid - 32];
    } else if (tid >= 64 && tid < 96) {
      int r = tid - 64;   // zero-filled rows contribute nothing
      for (int c = 0; c < D; ++c) l1 += fabs((double)sr[r][c] - (double)sg[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int e = tid + q * 256;
    if (e < D * D) {
      atomicAdd(acc + acc_mom_g(D) + e, mg[q]);
      atomicAdd(acc + acc_mom_r(D) + e, mr[q]);
    }
  }
  if (tid < D) atomicAdd(acc + acc_sum_g(D) + tid, colsum);
  else if (tid >= 32 && tid < 32 + D) atomicAdd(acc + acc_sum_r(D) + (tid - 32), colsum);
  else if (tid >= 64 && tid < 96) {
    l1 = s2ag_warp_sum_d(l1);
    if (tid == 64) {
      atomicAdd(acc + 1, l1);
      if (blockIdx.x == 0) atomicAdd(acc + 0, (double)N);
    }
  }
}

// Cyclic Jacobi on the symmetric matrix A (D x D, shared, one warp; lane k owns row / column k).  On return the
// diagonal of A holds the eigenvalues and, when V != nullptr, the columns of V the eigenvectors.
__device__ void jacobi_warp(double (*A)[kP], double (*V)[kP], int D, int lane) {
  if (V != nullptr && lane < D)
    for (int j = 0; j < D; ++j) V[lane][j] = (lane == j) ? 1.0 : 0.0;
  __syncwarp();
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0, dia = 0;
    if (lane < D)
      for (int j = 0; j < D; ++j) {
        double v = A[lane][j] * A[lane][j];
        if (j == lane) dia += v; else off += v;
      }
    off = s2ag_warp_sum_d(off); dia = s2ag_warp_sum_d(dia);
    if (!(off > 1e-30 * dia)) break;   // also leaves on NaN
    for (int p = 0; p < D - 1; ++p)
      for (int q = p + 1; q < D; ++q) {
        double apq = A[p][q];
        if (apq == 0.0) continue;   // warp-uniform: every lane reads the same element
        double app = A[p][p], aqq = A[q][q];
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        double akp = 0, akq = 0, vkp = 0, vkq = 0;
        if (lane < D) {
          akp = A[lane][p]; akq = A[lane][q];
          if (V != nullptr) { vkp = V[lane][p]; vkq = V[lane][q]; }
        }
        __syncwarp();
        if (lane < D) {
          if (lane != p && lane != q) {
            double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
            A[lane][p] = np_; A[p][lane] = np_;
            A[lane][q] = nq_; A[q][lane] = nq_;
          } else if (lane == p) {
            A[p][p] = app - t * apq; A[p][q] = 0.0;
          } else {
            A[q][q] = aqq + t * apq; A[q][p] = 0.0;
          }
          if (V != nullptr) { V[lane][p] = c * vkp - s * vkq; V[lane][q] = s * vkp + c * vkq; }
        }
        __syncwarp();
      }
  }
}

// one warp.  mode 0: moments from the accumulation buffer (np.mean / np.cov(rowvar=False), i.e. the unbiased
// covariance); mode 1: mu1 / sigma1 / mu2 / sigma2 given (calculate_frechet_distance's own arguments).
__global__ void __launch_bounds__(32) fgd_scores_kernel(const double* __restrict__ acc, const double* __restrict__ mu1_in,
                                                         const double* __restrict__ s1_in,
                                                         const double* __restrict__ mu2_in,
                                                         const double* __restrict__ s2_in, int D, int mode,
                                                         double* __restrict__ out) {
  __shared__ double S1[kMaxD][kP], S2[kMaxD][kP], V[kMaxD][kP], Tm[kMaxD][kP];
  const int lane = threadIdx.x;
  double n = 0, dmu = 0, tr = 0;
  if (mode == 0) {
    n = acc[0];
    if (lane < D) {
      double mg = acc[acc_sum_g(D) + lane] / n, mr = acc[acc_sum_r(D) + lane] / n;
      dmu = (mg - mr) * (mg - mr);
      for (int j = 0; j < D; ++j) {
        double mgj = acc[acc_sum_g(D) + j] / n, mrj = acc[acc_sum_r(D) + j] / n;
        S1[lane][j] = (acc[acc_mom_g(D) + (long)lane * D + j] - n * mg * mgj) / (n - 1.0);
        S2[lane][j] = (acc[acc_mom_r(D) + (long)lane * D + j] - n * mr * mrj) / (n - 1.0);
      }
    }
  } else if (lane < D) {
    dmu = (mu1_in[lane] - mu2_in[lane]) * (mu1_in[lane] - mu2_in[lane]);
    for (int j = 0; j < D; ++j) {
      S1[lane][j] = s1_in[(long)lane * D + j];
      S2[lane][j] = s2_in[(long)lane * D + j];
    }
  }
  __syncwarp();
  if (lane < D) tr = S1[lane][lane] + S2[lane][lane];
  // symmetrise (the moment route is symmetric up to rounding, caller-given matrices need not be bit-symmetric)
  if (lane < D)
    for (int j = 0; j < D; ++j) Tm[lane][j] = 0.5 * (S1[lane][j] + S1[j][lane]);
  __syncwarp();
  jacobi_warp(Tm, V, D, lane);
  // S1 <- S1^(1/2) = V diag(sqrt(max(lambda, 0))) V^T
  if (lane < D)
    for (int j = 0; j < D; ++j) {
      double a = 0;
      for (int k = 0; k < D; ++k) {
        double lam = Tm[k][k];
        a += V[lane][k] * sqrt(lam > 0 ? lam : 0.0) * V[j][k];
      }
      S1[lane][j] = a;
    }
  __syncwarp();
  // V <- S1^(1/2) S2 ; Tm <- V S1^(1/2), symmetrised
  if (lane < D)
    for (int j = 0; j < D; ++j) {
      double a = 0;
      for (int k = 0; k < D; ++k) a += S1[lane][k] * 0.5 * (S2[k][j] + S2[j][k]);
      V[lane][j] = a;
    }
  __syncwarp();
  if (lane < D)
    for (int j = 0; j < D; ++j) {
      double a = 0;
      for (int k = 0; k < D; ++k) a += V[lane][k] * S1[k][j];
      S2[lane][j] = a;
    }
  __syncwarp();
  if (lane < D)
    for (int j = 0; j < D; ++j) Tm[lane][j] = 0.5 * (S2[lane][j] + S2[j][lane]);
  __syncwarp();
  jacobi_warp(Tm, nullptr, D, lane);
  double rt = 0;
  if (lane < D) { double lam = Tm[lane][lane]; rt = lam > 0 ? sqrt(lam) : (lam == lam ? 0.0 : lam); }
  dmu = s2ag_warp_sum_d(dmu); tr = s2ag_warp_sum_d(tr); rt = s2ag_warp_sum_d(rt);
  if (lane == 0) {
    out[0] = dmu + tr - 2.0 * rt;
    out[1] = mode == 0 ? acc[1] / n : 0.0;
  }
}

}  // namespace

extern "C" long s2ag_fgd_acc_doubles(int D) { return 2 + 2 * (long)D + 2 * (long)D * D; }

extern "C" int s2ag_fgd_accumulate(const float* gen_feat, long ld_gen, const float* real_feat, long ld_real, int N, int D,
                                   double* acc, void* stream) {
  S2AG_CHECK_ARG(gen_feat && real_feat && acc && N >= 0 && D > 0 && D <= kMaxD && ld_gen >= D && ld_real >= D);
  if (N == 0) return S2AG_OK;
  int blocks = s2ag_cdiv(N, kRows);
  if (blocks > s2ag_sm_count()) blocks = s2ag_sm_count();
  auto k = &fgd_accumulate_kernel;
  S2AG_LAUNCH(k, blocks, 256, 0, stream, gen_feat, ld_gen, real_feat, ld_real, N, D, acc);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_fgd_scores(const double* acc, int D, double* out2, void* stream) {
  S2AG_CHECK_ARG(acc && out2 && D > 0 && D <= kMaxD);
  auto k = &fgd_scores_kernel;
  S2AG_LAUNCH(k, 1, 32, 0, stream, acc, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr,
              (const double*)nullptr, D, 0, out2);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_frechet_distance(const double* mu1, const double* sigma1, const double* mu2, const double* sigma2,
                                     int D, double* out2, void* stream) {
  S2AG_CHECK_ARG(mu1 && sigma1 && mu2 && sigma2 && out2 && D > 0 && D <= kMaxD);
  auto k = &fgd_scores_kernel;
  S2AG_LAUNCH(k, 1, 32, 0, stream, (const double*)nullptr, mu1, sigma1, mu2, sigma2, D, 1, out2);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
