// Fused WavEncoder forward for sm_100a: the raw-audio Conv1d stack of the frozen tri-modal baseline
// (net/multimodal_context_net_v2.py:14-33: Conv1d(1,16,15,s5,p1600) BN LeakyReLU(0.3) Conv1d(16,32,15,s6) BN LeakyReLU
// Conv1d(32,64,15,s6) BN LeakyReLU Conv1d(64,32,15,s6); [B, L] -> [B, 34, 32]), BatchNorm in train mode (batch
// statistics; the reference never calls .eval() on the baseline during training, processor_v2.py:961-962) or eval mode.
//
// What the unfused chain paid: conv1's [B,7891,16] fp32 output (129 MB at 256 clips) was written, re-read for the
// statistics, re-read + rewritten by the normalisation, re-read by conv2 (and again for conv2 / conv3): ~18x the
// algorithmic bytes (149 420 B per clip = audio in + features out).  Here:
//   K1  wav_conv1_stats_kernel   audio -> per-channel shifted sums of conv1's output; the output itself is NEVER stored
//   K2  wav_conv_kernel<16,32,1> audio -> conv1 recomputed per tile (fp32 SIMT) -> BN1 + LeakyReLU -> bf16 hi/lo operand
//                                images in shared memory -> conv2 on tcgen05 -> raw conv2 output (L2-sized, 43 MB) + its sums
//   K3  wav_conv_kernel<32,64,0> raw conv2 -> BN2 + LeakyReLU while the operand is staged -> conv3 -> raw output + sums
//   K4  wav_conv_kernel<64,32,0> raw conv3 -> BN3 + LeakyReLU while staged -> conv4 -> features (row stride ldy: the
//                                GRU input buffer's column slice)
// The BatchNorm scale / shift are derived from the sums in each consumer's prologue (CTA 0 also updates the running
// statistics): no finalize launches, no normalised copy of any activation in HBM.
//
// Strided convolution as a shifted-window tcgen05 contraction.  A stride-6 filter does not overlap as a row shift of
// one image (umma_conv.cu), but it does per PHASE: with tap k = 6q + r,
//     y[o] = sum_r sum_q W[6q + r] . X_r[o + q],        X_r[j] = x[6j + r]   (the phase-r subsampled image),
// i.e. 6 stride-1 convolutions with 3 (r < 3) or 2 taps over the phase images.  A tile of 128 anchors stages every input
// pixel it needs exactly ONCE into its phase image ([phase][8-channel chunk][row][16 B], K-major UMMA layout, 130 rows);
// tap k is then one tcgen05.mma per 16 input channels whose A descriptor starts q rows into phase r.  Anchors run over
// the concatenation of all clips with a per-clip pitch of ceil(Lin / 6) rows, so tiles cross clip boundaries without
// padding waste (anchors o >= Lout of a clip are discarded).
#include "s2ag.h"
#include "gemm_umma.cuh"

namespace s2ag {
namespace wav {

using namespace s2ag::umma;

constexpr int KT = 15;                 // filter taps of every layer
constexpr int S1 = 5, PAD1 = 1600;     // conv1 stride / padding
constexpr int C1 = 16;                 // conv1 output channels
constexpr int S = 6;                   // stride of conv2..4 = number of phase images
constexpr int QH = (KT - 1) / S;       // halo rows per phase image (2)
constexpr int TM = 128;                // anchors per tile
constexpr int R = TM + QH;             // staged rows per phase image
constexpr int THREADS = 384, WORK = 256, NISS = 4, HDR = 1024;
constexpr int AUD_FLOATS = 4096;       // audio span of one tile: <= 30 * R + 2 * 16 samples

struct BnIn {                          // BatchNorm between the producing convolution and this one
  const double* sums;                  // [2*C] shifted sums of the producer's raw output relative to its bias (training)
  const float* bias_prev;              // the producer's bias (what the sums are relative to)
  const float* gamma; const float* beta; float* rmean; float* rvar;
  long count; int training; float momentum, eps, slope;
};

struct Params {
  const float* x;                      // raw producer output [B, Lin, CIN] (FUSE1: the audio [B, L])
  int B, L, Lin, Lo, pitch;            // L: audio samples (FUSE1); Lin/Lo: input / output pixels per clip
  const float* w1; const float* b1;    // FUSE1: conv1 weight [16][1][15], bias
  const float* w; const float* bias;   // this layer's weight [COUT][CIN][15], bias
  float* y; long ldy;                  // raw (pre-BatchNorm) output rows, or the final features
  double* sums_out;                    // [2*COUT] or NULL
  BnIn bn;
  int tiles; long total_rows;          // B * pitch
  int x3;
};

__device__ __forceinline__ void pack8w(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * p], v[2 * p + 1]);
    h[p] = *reinterpret_cast<const uint32_t*>(&hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * p] - __low2float(hh), v[2 * p + 1] - __high2float(hh));
    l[p] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void tmem_ld8w(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}

// BatchNorm scale / shift of channel c from the producer's shifted sums (training) or the running statistics
__device__ __forceinline__ void bn_scale_shift(const BnIn& bn, int c, int C, bool update_running, float& scale, float& shift) {
  float mean, invstd;
  if (bn.training) {
    const double e1 = bn.sums[c] / (double)bn.count, e2 = bn.sums[C + c] / (double)bn.count;
    double var = e2 - e1 * e1; if (var < 0.0) var = 0.0;
    mean = (float)((double)bn.bias_prev[c] + e1);
    invstd = (float)(1.0 / sqrt(var + (double)bn.eps));
    if (update_running && bn.rmean) {
      const float unbiased = (float)(bn.count > 1 ? var * (double)bn.count / (double)(bn.count - 1) : var);
      bn.rmean[c] = (1.f - bn.momentum) * bn.rmean[c] + bn.momentum * mean;
      bn.rvar[c] = (1.f - bn.momentum) * bn.rvar[c] + bn.momentum * unbiased;
    }
  } else {
    mean = bn.rmean[c];
    invstd = 1.f / sqrtf(bn.rvar[c] + bn.eps);
  }
  const float g = bn.gamma ? bn.gamma[c] : 1.f, b = bn.beta ? bn.beta[c] : 0.f;
  scale = g * invstd;
  shift = b - mean * g * invstd;
}

// ---------------------------------------------------------------------------------------------------------------- K1
// sums[c] += sum_p (conv1(audio)[p, c] - b1[c]),  sums[16 + c] += sum_p (.)^2  over all pixels of all clips.
// A block stages the audio span of 1024 consecutive output pixels of one clip (5130 samples) and every thread computes
// 4 pixels x 16 channels (weights by 128-bit broadcast loads); the conv1 output never leaves registers.
constexpr int K1_PIX = 1024, K1_SPAN = S1 * (K1_PIX - 1) + KT + 1;
__global__ void __launch_bounds__(256) wav_conv1_stats_kernel(const float* __restrict__ audio, int B, int L, int L1,
                                                              const float* __restrict__ w1, double* __restrict__ sums,
                                                              int items_per_clip) {
  __shared__ __align__(16) float ws[KT][C1];
  __shared__ float aud[K1_SPAN];
  __shared__ double red[2][C1];
  const int tid = threadIdx.x;
  for (int i = tid; i < KT * C1; i += 256) ws[i / C1][i % C1] = __ldg(w1 + (i % C1) * KT + i / C1);
  if (tid < 2 * C1) red[tid / C1][tid % C1] = 0.0;
  float s1[C1], s2[C1];
#pragma unroll
  for (int c = 0; c < C1; ++c) s1[c] = s2[c] = 0.f;
  const int items = B * items_per_clip;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int clip = item / items_per_clip, p0 = (item - clip * items_per_clip) * K1_PIX;
    const int npix = min(K1_PIX, L1 - p0);
    const int s_lo = S1 * p0 - PAD1, span = S1 * (npix - 1) + KT;
    const float* src = audio + (long)clip * L;
    __syncthreads();
    for (int i = tid; i < span; i += 256) {
      const int s = s_lo + i;
      aud[i] = (s >= 0 && s < L) ? __ldg(src + s) : 0.f;
    }
    __syncthreads();
    float acc[4][C1];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < C1; ++c) acc[i][c] = 0.f;
    int off[4]; bool ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int p = tid + 256 * i; ok[i] = p < npix; off[i] = ok[i] ? S1 * p : 0; }
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      float wv[C1];
#pragma unroll
      for (int c4 = 0; c4 < C1 / 4; ++c4) {
        const float4 q = *reinterpret_cast<const float4*>(&ws[t][4 * c4]);
        wv[4 * c4] = q.x; wv[4 * c4 + 1] = q.y; wv[4 * c4 + 2] = q.z; wv[4 * c4 + 3] = q.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xv = aud[off[i] + t];
#pragma unroll
        for (int c = 0; c < C1; ++c) acc[i][c] = fmaf(xv, wv[c], acc[i][c]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) {
#pragma unroll
        for (int c = 0; c < C1; ++c) { s1[c] += acc[i][c]; s2[c] = fmaf(acc[i][c], acc[i][c], s2[c]); }
      }
  }
#pragma unroll
  for (int c = 0; c < C1; ++c) {
    const float a = s2ag_warp_sum(s1[c]), b = s2ag_warp_sum(s2[c]);
    if ((tid & 31) == 0) { atomicAdd(&red[0][c], (double)a); atomicAdd(&red[1][c], (double)b); }
  }
  __syncthreads();
  if (tid < 2 * C1) atomicAdd(sums + tid, red[tid / C1][tid % C1]);
}

// ------------------------------------------------------------------------------------------------------------ K2..K4
template <int CIN, int COUT, bool FUSE1>
struct Geo {
  static constexpr int CPASS = CIN < 32 ? CIN : 32;      // input channels staged per pass
  static constexpr int NPASS = CIN / CPASS;
  static constexpr int KC = CPASS / 8;                   // 8-channel chunks per pass
  static constexpr int PST = KC * R + 1;                 // phase stride in 16-byte units (+1: conflict-free staging)
  static constexpr int A_PLANE = S * PST * 16;
  static constexpr int W_PLANE = KT * KC * COUT * 16;
  static constexpr int OFF_SC = HDR;                                  // scale[CIN], shift[CIN]
  static constexpr int OFF_W = OFF_SC + 2 * 64 * 4;
  static constexpr int OFF_A = OFF_W + 2 * W_PLANE;
  static constexpr int OFF_AUD = OFF_A + 2 * A_PLANE;                 // FUSE1: audio span + conv1 weights
  static constexpr int OFF_W1 = OFF_AUD + (FUSE1 ? AUD_FLOATS * 4 : 0);
  static constexpr int OFF_RED = OFF_W1 + (FUSE1 ? (KT * C1 + C1) * 4 : 0);
  static constexpr int SMEM = OFF_RED + 2 * COUT * 8;
  static constexpr uint32_t NCOLS = NISS * COUT <= 32 ? 32 : NISS * COUT <= 64 ? 64 : NISS * COUT <= 128 ? 128 : NISS * COUT <= 256 ? 256 : 512;
  static_assert(CIN % CPASS == 0 && CPASS % 16 == 0 && COUT % 16 == 0 && NISS * COUT <= 512, "unsupported channel counts");
  static_assert(!FUSE1 || CIN == C1, "the fused first layer produces 16 channels");
};

template <int CIN, int COUT, bool FUSE1>
__global__ void __launch_bounds__(THREADS, 1) wav_conv_kernel(Params p) {
  using G = Geo<CIN, COUT, FUSE1>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t mma_bar = sbase;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 16);
  float* sc = reinterpret_cast<float*>(smem + G::OFF_SC);
  float* sh = sc + 64;
  unsigned char* w_hi = smem + G::OFF_W;
  unsigned char* w_lo = w_hi + G::W_PLANE;
  unsigned char* a_hi = smem + G::OFF_A;
  unsigned char* a_lo = a_hi + G::A_PLANE;
  float* aud = reinterpret_cast<float*>(smem + G::OFF_AUD);
  float* w1s = reinterpret_cast<float*>(smem + G::OFF_W1);     // [KT][16] then bias[16]
  double* red = reinterpret_cast<double*>(smem + G::OFF_RED);  // [2][COUT]

  if (tid == 0) {
    mbar_init(mma_bar, NISS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 16, G::NCOLS);
  // ---- prologue: BatchNorm of the input as scale / shift (CTA 0 updates the running statistics once)
  if (tid < CIN) {
    float a, b;
    bn_scale_shift(p.bn, tid, CIN, blockIdx.x == 0, a, b);
    sc[tid] = a; sh[tid] = b;
  }
  for (int i = tid; i < 2 * COUT; i += THREADS) red[i] = 0.0;
  if (FUSE1) {
    for (int i = tid; i < KT * C1; i += THREADS) w1s[i] = __ldg(p.w1 + (i % C1) * KT + i / C1);
    if (tid < C1) w1s[KT * C1 + tid] = __ldg(p.b1 + tid);
  }
  auto stage_w = [&](int pass) {
    // element (tap t, out channel n, in channel pass*CPASS + kc*8 + i) of the reference weight [COUT][CIN][KT]
    for (int idx = tid; idx < KT * G::KC * COUT; idx += THREADS) {
      const int n = idx % COUT; const int kc = (idx / COUT) % G::KC; const int t = idx / (COUT * G::KC);
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(p.w + ((long)n * CIN + pass * G::CPASS + kc * 8 + i) * KT + t);
      uint4 hi, lo;
      pack8w(v, hi, lo);
      *reinterpret_cast<uint4*>(w_hi + idx * 16) = hi;
      *reinterpret_cast<uint4*>(w_lo + idx * 16) = lo;
    }
  };
  if (G::NPASS == 1) stage_w(0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool vec_dst = (p.ldy & 7) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0;
  constexpr int CH = COUT / 2;          // output channels per epilogue thread
  float s1[CH], s2[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) s1[c] = s2[c] = 0.f;

  uint32_t parity = 0;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long a0 = (long)tile * TM;
    for (int pass = 0; pass < G::NPASS; ++pass, parity ^= 1u) {
      if (G::NPASS > 1) stage_w(pass);
      if (warp < 8) {
        if (FUSE1) {
          // ---- audio span of the tile's rows: at most two clips (pitch > R)
          const int c0 = (int)(a0 / p.pitch), j0 = (int)(a0 - (long)c0 * p.pitch);
          const int n0 = min(R, p.pitch - j0);           // rows of the first clip
          const int seg1 = S1 * S * n0 + 16;             // smem offset of the second clip's span
          for (int i = tid; i < S1 * S * n0 + KT; i += WORK) {
            const int s = S1 * S * j0 - PAD1 + i;
            aud[i] = (c0 < p.B && s >= 0 && s < p.L) ? __ldg(p.x + (long)c0 * p.L + s) : 0.f;
          }
          if (n0 < R) {
            const int n1 = R - n0;
            for (int i = tid; i < S1 * S * n1 + KT; i += WORK) {
              const int s = i - PAD1;
              aud[seg1 + i] = (c0 + 1 < p.B && s >= 0 && s < p.L) ? __ldg(p.x + (long)(c0 + 1) * p.L + s) : 0.f;
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          // ---- conv1 (fp32 SIMT) + BN1 + LeakyReLU -> operand images; thread = 3 pixels x 16 channels
          for (int base = 0; base < R * S; base += 3 * WORK) {
            if (base + tid >= R * S) break;
            int off[3], dst[3]; bool ok[3], real[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int idx = base + tid + WORK * i;       // idx = r * R + row: consecutive threads, consecutive rows
              ok[i] = idx < R * S;
              const int r = ok[i] ? idx / R : 0, row = ok[i] ? idx - r * R : 0;
              const bool second = row >= n0;
              const int j = second ? row - n0 : j0 + row;
              const int pix = S * j + r;
              real[i] = ok[i] && (c0 + (second ? 1 : 0)) < p.B && pix < p.Lin;
              off[i] = (second ? seg1 + S1 * pix : S1 * (pix - S * j0));
              if (!real[i]) off[i] = 0;
              dst[i] = (r * G::PST + row) * 16;
            }
            float acc[3][C1];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int c = 0; c < C1; ++c) acc[i][c] = 0.f;
#pragma unroll
            for (int t = 0; t < KT; ++t) {
              float wv[C1];
#pragma unroll
              for (int c4 = 0; c4 < C1 / 4; ++c4) {
                const float4 q = *reinterpret_cast<const float4*>(w1s + t * C1 + 4 * c4);
                wv[4 * c4] = q.x; wv[4 * c4 + 1] = q.y; wv[4 * c4 + 2] = q.z; wv[4 * c4 + 3] = q.w;
              }
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const float xv = aud[off[i] + t];
#pragma unroll
                for (int c = 0; c < C1; ++c) acc[i][c] = fmaf(xv, wv[c], acc[i][c]);
              }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              if (!ok[i]) continue;
#pragma unroll
              for (int kc = 0; kc < 2; ++kc) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const int c = kc * 8 + e;
                  const float tv = fmaf(acc[i][c] + w1s[KT * C1 + c], sc[c], sh[c]);
                  v[e] = real[i] ? (tv > 0.f ? tv : tv * p.bn.slope) : 0.f;
                }
                uint4 hi, lo;
                pack8w(v, hi, lo);
                *reinterpret_cast<uint4*>(a_hi + dst[i] + kc * R * 16) = hi;
                if (p.x3) *reinterpret_cast<uint4*>(a_lo + dst[i] + kc * R * 16) = lo;
              }
            }
          }
        } else {
          // ---- raw producer output -> BN + LeakyReLU -> operand images; item = (local pixel, 8-channel chunk),
          //      consecutive threads read consecutive 32-byte segments
          for (int it = tid; it < R * S * G::KC; it += WORK) {
            const int kc = it % G::KC, lp = it / G::KC;
            const int row = lp / S, r = lp - row * S;
            const long g = a0 + row;
            const int c = (int)(g / p.pitch), j = (int)(g - (long)c * p.pitch);
            const int pix = S * j + r;
            float v[8];
            if (c < p.B && pix < p.Lin) {
              const int ch = pass * G::CPASS + kc * 8;
              const float4* src = reinterpret_cast<const float4*>(p.x + ((long)c * p.Lin + pix) * CIN + ch);
              const float4 a = __ldg(src), b = __ldg(src + 1);
              v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float tv = fmaf(v[e], sc[ch + e], sh[ch + e]);
                v[e] = tv > 0.f ? tv : tv * p.bn.slope;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = 0.f;
            }
            uint4 hi, lo;
            pack8w(v, hi, lo);
            const int dst = (r * G::PST + kc * R + row) * 16;
            *reinterpret_cast<uint4*>(a_hi + dst) = hi;
            if (p.x3) *reinterpret_cast<uint4*>(a_lo + dst) = lo;
          }
        }
      }
      fence_async_smem();
      __syncthreads();
      if (warp_u >= 8 && elect_one()) {
        // ---- issuer `iss`: taps iss, iss + NISS, ... of this pass into accumulator iss
        const int iss = warp_u - 8;
        tc_fence_after();
        const uint32_t idesc = make_idesc(COUT);
        const uint32_t sa = smem_u32(a_hi), sw = smem_u32(w_hi);
        const uint32_t d = tmem_base + (uint32_t)(iss * COUT);
        uint32_t cnt = pass;   // pass > 0 accumulates onto pass 0
        for (int t = iss; t < KT; t += NISS) {
          const int q = t / S, r = t - q * S;
#pragma unroll
          for (int k2 = 0; k2 < G::KC / 2; ++k2) {
            const uint32_t ah = sa + (uint32_t)((r * G::PST + 2 * k2 * R + q) * 16), al = ah + (uint32_t)G::A_PLANE;
            const uint32_t wh = sw + (uint32_t)(((t * G::KC + 2 * k2) * COUT) * 16), wl = wh + (uint32_t)G::W_PLANE;
            const uint64_t dah = make_desc(ah, R * 16, 128), dwh = make_desc(wh, COUT * 16, 128);
            if (p.x3) {
              mma_bf16(d, make_desc(al, R * 16, 128), dwh, idesc, cnt ? 1u : 0u); ++cnt;
              mma_bf16(d, dah, make_desc(wl, COUT * 16, 128), idesc, 1u); ++cnt;
            }
            mma_bf16(d, dah, dwh, idesc, cnt ? 1u : 0u); ++cnt;
          }
        }
        mma_commit(mma_bar);
      }
      mbar_wait(mma_bar, parity);
      tc_fence_after();
      if (pass + 1 < G::NPASS) { tc_fence_before(); __syncthreads(); }   // operands of the next pass overwrite these
    }
    if (warp < 8) {
      // ---- epilogue: thread = anchor, warpgroup = half of the output channels
      const int row = (warp & 3) * 32 + lane;
      const long a = a0 + row;
      const int c = (int)(a / p.pitch), o = (int)(a - (long)c * p.pitch);
      const bool ok = c < p.B && o < p.Lo;
      float* dst = p.y + ((long)c * p.Lo + o) * p.ldy;
      const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      const int c_beg = (warp >> 2) * CH;
#pragma unroll
      for (int cg = 0; cg < CH / 8; ++cg) {
        const int c0 = c_beg + cg * 8;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int ai = 0; ai < NISS; ++ai) {
          float v[8];
          tmem_ld8w(t_lane + (uint32_t)(ai * COUT + c0), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
        if (ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { s1[cg * 8 + i] += acc[i]; s2[cg * 8 + i] = fmaf(acc[i], acc[i], s2[cg * 8 + i]); }
          float val[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) val[i] = acc[i] + __ldg(p.bias + c0 + i);
          if (vec_dst) {
            asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"f"(val[0]), "f"(val[1]), "f"(val[2]),
                         "f"(val[3]), "f"(val[4]), "f"(val[5]), "f"(val[6]), "f"(val[7]), "l"(dst + c0)
                         : "memory");
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[c0 + i] = val[i];
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  // ---- per-channel sums of this layer's raw output (relative to its bias) for the next BatchNorm
  if (p.sums_out) {
    if (warp < 8) {
      const int c_beg = (warp >> 2) * CH;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float a = s2ag_warp_sum(s1[c]), b = s2ag_warp_sum(s2[c]);
        if (lane == 0) { atomicAdd(&red[c_beg + c], (double)a); atomicAdd(&red[COUT + c_beg + c], (double)b); }
      }
    }
    __syncthreads();
    for (int i = tid; i < 2 * COUT; i += THREADS) atomicAdd(p.sums_out + i, red[i]);
  }
  if (warp == 0) tmem_dealloc(tmem_base, G::NCOLS);
}

static inline int conv_len(int L, int k, int s, int pad) { return (L + 2 * pad - k) / s + 1; }

template <int CIN, int COUT, bool FUSE1>
static int launch_layer(Params& p, int sms, void* stream) {
  using G = Geo<CIN, COUT, FUSE1>;
  auto kfn = &wav_conv_kernel<CIN, COUT, FUSE1>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM) != cudaSuccess) return -1;
    attr_set = true;
  }
  p.total_rows = (long)p.B * p.pitch;
  p.tiles = (int)((p.total_rows + TM - 1) / TM);
  p.x3 = umma::g_precision == 0 ? 1 : 0;
  int grid = sms < p.tiles ? sms : p.tiles;
  S2AG_LAUNCH(kfn, grid, THREADS, G::SMEM, stream, p);
  return 0;
}

}  // namespace wav
}  // namespace s2ag

using namespace s2ag::wav;

extern "C" long s2ag_wavencoder_ws_floats(int B, int L) {
  if (B <= 0 || L <= 0) return 0;
  const int L1 = conv_len(L, KT, S1, PAD1);
  if (L1 < KT) return 0;
  const int L2 = conv_len(L1, KT, S, 0);
  if (L2 < KT) return 0;
  const int L3 = conv_len(L2, KT, S, 0);
  return (long)B * ((long)L2 * 32 + (long)L3 * 64) + 2 * (2 * 16 + 2 * 32 + 2 * 64) + 64;
}

extern "C" int s2ag_wavencoder_fwd(const float* audio, int B, int L, const float* const* conv_w,
                                   const float* const* conv_b, const float* const* bn_gamma,
                                   const float* const* bn_beta, float* const* bn_rmean, float* const* bn_rvar,
                                   int training, float momentum, float eps, float slope, float* y, long ldy, float* ws,
                                   void* stream) {
  S2AG_CHECK_ARG(audio && conv_w && conv_b && bn_gamma && bn_beta && bn_rmean && bn_rvar && y && ws && B > 0 && L > 0);
  for (int i = 0; i < 4; ++i) S2AG_CHECK_ARG(conv_w[i] && conv_b[i]);
  for (int i = 0; i < 3; ++i) S2AG_CHECK_ARG(bn_rmean[i] && bn_rvar[i]);
  const int L1 = conv_len(L, KT, S1, PAD1);
  S2AG_CHECK_ARG(L1 >= KT);
  const int L2 = conv_len(L1, KT, S, 0);
  S2AG_CHECK_ARG(L2 >= KT);
  const int L3 = conv_len(L2, KT, S, 0);
  S2AG_CHECK_ARG(L3 >= KT);
  const int L4 = conv_len(L3, KT, S, 0);
  S2AG_CHECK_ARG(ldy >= 32 && (long)B * L1 < (1L << 31));
  S2AG_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 31) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  static int sms = 0;
  if (!sms) {
    int dev = 0; cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  // workspace: raw conv2 output, raw conv3 output, sums (doubles) of conv1 / conv2 / conv3
  float* y2 = ws;
  float* y3 = y2 + (long)B * L2 * 32;
  double* sums = reinterpret_cast<double*>(y3 + (long)B * L3 * 64 + (((long)B * L3 * 64) & 1));
  S2AG_CHECK_ARG((reinterpret_cast<uintptr_t>(sums) & 7) == 0);
  double* sums1 = sums; double* sums2 = sums + 2 * 16; double* sums3 = sums2 + 2 * 32;
  if (training) {
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (16 + 32 + 64), st);
    const int ipc = (L1 + K1_PIX - 1) / K1_PIX;
    int grid = B * ipc; if (grid > sms * 4) grid = sms * 4;
    auto k1 = &wav_conv1_stats_kernel;
    S2AG_LAUNCH(k1, grid, 256, 0, stream, audio, B, L, L1, conv_w[0], sums1, ipc);
  }
  auto bn_in = [&](int i, double* s, long count) {
    BnIn b;
    b.sums = s; b.bias_prev = conv_b[i]; b.gamma = bn_gamma[i]; b.beta = bn_beta[i]; b.rmean = bn_rmean[i];
    b.rvar = bn_rvar[i]; b.count = count; b.training = training; b.momentum = momentum; b.eps = eps; b.slope = slope;
    return b;
  };
  Params p;
  // conv1 (recomputed) + BN1 -> conv2
  p.x = audio; p.B = B; p.L = L; p.Lin = L1; p.Lo = L2; p.pitch = (L1 + S - 1) / S;
  p.w1 = conv_w[0]; p.b1 = conv_b[0]; p.w = conv_w[1]; p.bias = conv_b[1]; p.y = y2; p.ldy = 32;
  p.sums_out = training ? sums2 : nullptr; p.bn = bn_in(0, sums1, (long)B * L1);
  S2AG_CHECK_ARG(p.pitch > R);
  if (launch_layer<16, 32, true>(p, sms, stream)) { s2ag_set_error("wavencoder: shared memory attribute"); return S2AG_ERR_LAUNCH; }
  // BN2 -> conv3
  p.x = y2; p.Lin = L2; p.Lo = L3; p.pitch = (L2 + S - 1) / S; p.w1 = nullptr; p.b1 = nullptr;
  p.w = conv_w[2]; p.bias = conv_b[2]; p.y = y3; p.ldy = 64; p.sums_out = training ? sums3 : nullptr;
  p.bn = bn_in(1, sums2, (long)B * L2);
  if (launch_layer<32, 64, false>(p, sms, stream)) { s2ag_set_error("wavencoder: shared memory attribute"); return S2AG_ERR_LAUNCH; }
  // BN3 -> conv4 -> features
  p.x = y3; p.Lin = L3; p.Lo = L4; p.pitch = (L3 + S - 1) / S; p.w = conv_w[3]; p.bias = conv_b[3]; p.y = y; p.ldy = ldy;
  p.sums_out = nullptr; p.bn = bn_in(2, sums3, (long)B * L3);
  if (launch_layer<64, 32, false>(p, sms, stream)) { s2ag_set_error("wavencoder: shared memory attribute"); return S2AG_ERR_LAUNCH; }
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
