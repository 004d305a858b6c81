// Fused WavEncoder forward for sm_100a: the raw-audio Conv1d stack of the frozen tri-modal baseline
// (net/multimodal_context_net_v2.py:14-33: Conv1d(1,16,15,s5,p1600) BN LeakyReLU(0.3) Conv1d(16,32,15,s6) BN LeakyReLU
// Conv1d(32,64,15,s6) BN LeakyReLU Conv1d(64,32,15,s6); [B, L] -> [B, 34, 32]), BatchNorm in train mode (batch
// statistics; the reference never calls .eval() on the baseline during training, processor_v2.py:961-962) or eval mode.
//
// What the unfused chain paid: conv1's [B,7891,16] fp32 output (129 MB at 256 clips) was written, re-read for the
// statistics, re-read + rewritten by the normalisation, re-read by conv2 (and again for conv2 / conv3): ~18x the
// algorithmic bytes (149 420 B per clip = audio in + features out).  Here:
//   K1  wav_prep_kernel          audio -> per-channel shifted sums of conv1's output (the output itself is NEVER stored;
//                                audio spans arrive by double-buffered cp.async); extra blocks pack the weights of
//                                conv2..4 once into bf16 hi/lo tensor-core operand images
//   K2  wav_conv_kernel<16,32,1> audio -> conv1 recomputed per tile (fp32 SIMT) -> BN1 + LeakyReLU -> operand images in
//                                shared memory -> conv2 on tcgen05 -> raw conv2 output (L2-sized, 43 MB) + its sums
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
//
// One persistent CTA per SM, warp-specialised, mbarrier pipeline over items = (tile, 16-channel pass):
//   warps 0-13  producers: stage the operand images of item n into buffer n&1 (waits "empty": the MMAs of item n-2);
//               K3/K4: two groups of 7 warps alternate items so one group's loads fly while the other converts
//   warp  14    issuer: waits "full", issues the 15 taps x 3 (bf16x3) tcgen05.mma of the item into TMEM accumulator
//               tile&1, commits "empty" and, after the tile's last pass, "tmem full"
//   warps 15-22 epilogue: TMEM -> registers (bias, statistics) -> global, then "tmem empty"; overlaps the next tile
// The whole weight tensor of the layer stays in shared memory (one TMA bulk copy of the pre-packed image per CTA).
#include "s2ag.h"
#include "gemm_umma.cuh"
#include "gemm_umma_packed.cuh"

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
constexpr int RC = 132;                // chunk stride in rows (>= R, = 4 mod 8: conflict-free staging stores)
constexpr int PST = 2 * RC + 1;        // phase stride in 16-byte units (= 1 mod 8)
constexpr int A_PLANE = S * PST * 16;  // one bf16 plane of one operand buffer (16 channels)
constexpr int A_BUF = 2 * A_PLANE;     // hi + lo
constexpr int NPW = 14;               // staging warps (the fused first layer: one group; the others: two groups of 7)
constexpr int NPROD = NPW * 32, THREADS = (NPW + 1 + 8) * 32, HDR = 1024;
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
  const unsigned char* wpk;            // this layer's packed weight image [plane][tap][CIN/8][COUT][16 B]
  const float* bias;
  float* y; long ldy;                  // raw (pre-BatchNorm) output rows, or the final features
  double* sums_out;                    // [2*COUT] or NULL
  BnIn bn;
  int tiles; int total_rows;           // B * pitch
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
// 4-byte asynchronous copy global -> shared; !valid: zero fill (src is not read)
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
// blocks [0, stat_blocks): sums[c] += sum_p (conv1(audio)[p, c] - b1[c]), sums[16 + c] += sum_p (.)^2 over all pixels of
// all clips.  A block takes items of 1024 consecutive output pixels of one clip; the item's audio span (5130 samples)
// arrives by 4-byte cp.async into a double buffer while the previous item is computed; every thread computes 4 pixels x
// 16 channels (weights by 128-bit broadcast loads).  The conv1 output never leaves registers.
// blocks [stat_blocks, ...): pack the weights [COUT][CIN][15] of conv2..4 into the tensor-core operand images
// [plane hi|lo][tap][CIN/8][COUT][16 B] that the consumer CTAs fetch with one TMA bulk copy.
constexpr int K1_PIX = 1024, K1_SPAN = S1 * (K1_PIX - 1) + KT + 1;
struct PackJob { const float* w; unsigned char* dst; int cin, cout; };
struct PrepParams {
  const float* audio; int B, L, L1; const float* w1; double* sums; int items_per_clip, stat_blocks;
  PackJob job[3];
};

__global__ void __launch_bounds__(256) wav_prep_kernel(PrepParams p) {
  __shared__ __align__(16) float ws[KT][C1];
  __shared__ float aud[2][K1_SPAN];
  __shared__ double red[2][C1];
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= p.stat_blocks) {
    const int nb = gridDim.x - p.stat_blocks, b = blockIdx.x - p.stat_blocks;
    for (int l = 0; l < 3; ++l) {
      const PackJob j = p.job[l];
      const int kct = j.cin / 8, plane = KT * kct * j.cout * 16;
      for (int idx = b * 256 + tid; idx < KT * kct * j.cout; idx += nb * 256) {
        const int n = idx % j.cout; const int kc = (idx / j.cout) % kct; const int t = idx / (j.cout * kct);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(j.w + ((long)n * j.cin + kc * 8 + i) * KT + t);
        uint4 hi, lo;
        pack8w(v, hi, lo);
        *reinterpret_cast<uint4*>(j.dst + (long)idx * 16) = hi;
        *reinterpret_cast<uint4*>(j.dst + plane + (long)idx * 16) = lo;
      }
    }
    return;
  }
  for (int i = tid; i < KT * C1; i += 256) ws[i / C1][i % C1] = __ldg(p.w1 + (i % C1) * KT + i / C1);
  if (tid < 2 * C1) red[tid / C1][tid % C1] = 0.0;
  float s1[C1], s2[C1];
#pragma unroll
  for (int c = 0; c < C1; ++c) s1[c] = s2[c] = 0.f;
  const int items = p.B * p.items_per_clip;
  auto prefetch = [&](int item, int buf) {
    const int clip = item / p.items_per_clip, p0 = (item - clip * p.items_per_clip) * K1_PIX;
    const int npix = min(K1_PIX, p.L1 - p0);
    const int s_lo = S1 * p0 - PAD1, span = S1 * (npix - 1) + KT;
    const float* src = p.audio + (long)clip * p.L;
    const uint32_t dst = smem_u32(&aud[buf][0]);
    for (int i = tid; i < span; i += 256) {
      const int s = s_lo + i;
      const bool ok = s >= 0 && s < p.L;
      cp_async4(dst + 4u * i, src + (ok ? s : 0), ok);
    }
  };
  int item = blockIdx.x, buf = 0;
  if (item < items) prefetch(item, 0);
  cp_async_commit();
  for (; item < items; item += p.stat_blocks, buf ^= 1) {
    if (item + p.stat_blocks < items) prefetch(item + p.stat_blocks, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int clip = item / p.items_per_clip, p0 = (item - clip * p.items_per_clip) * K1_PIX;
    const int npix = min(K1_PIX, p.L1 - p0);
    const float* a = aud[buf];
    float2 acc[4][C1 / 2];   // packed fp32 pairs: FFMA2 (two FMAs per issue slot)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < C1 / 2; ++c) acc[i][c] = make_float2(0.f, 0.f);
    int off[4]; bool ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int px = tid + 256 * i; ok[i] = px < npix; off[i] = ok[i] ? S1 * px : 0; }
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      float2 wv[C1 / 2];
#pragma unroll
      for (int c4 = 0; c4 < C1 / 4; ++c4) {
        const float4 q = *reinterpret_cast<const float4*>(&ws[t][4 * c4]);
        wv[2 * c4] = make_float2(q.x, q.y); wv[2 * c4 + 1] = make_float2(q.z, q.w);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xv = a[off[i] + t];
        const float2 x2 = make_float2(xv, xv);
#pragma unroll
        for (int c = 0; c < C1 / 2; ++c) acc[i][c] = __ffma2_rn(x2, wv[c], acc[i][c]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) {
#pragma unroll
        for (int c = 0; c < C1 / 2; ++c) {
          s1[2 * c] += acc[i][c].x; s2[2 * c] = fmaf(acc[i][c].x, acc[i][c].x, s2[2 * c]);
          s1[2 * c + 1] += acc[i][c].y; s2[2 * c + 1] = fmaf(acc[i][c].y, acc[i][c].y, s2[2 * c + 1]);
        }
      }
    __syncthreads();   // aud[buf] is overwritten by the prefetch of the next iteration
  }
  cp_async_wait<0>();
#pragma unroll
  for (int c = 0; c < C1; ++c) {
    const float a = s2ag_warp_sum(s1[c]), b = s2ag_warp_sum(s2[c]);
    if ((tid & 31) == 0) { atomicAdd(&red[0][c], (double)a); atomicAdd(&red[1][c], (double)b); }
  }
  __syncthreads();
  if (tid < 2 * C1) atomicAdd(p.sums + tid, red[tid / C1][tid % C1]);
}

// ------------------------------------------------------------------------------------------------------------ K2..K4
template <int CIN, int COUT, bool FUSE1>
struct Geo {
  static constexpr int NPASS = CIN / 16;                 // 16 input channels per pipeline item
  static constexpr int KCT = CIN / 8;                    // 8-channel chunks of the whole weight image
  static constexpr int W_PLANE = KT * KCT * COUT * 16;
  static constexpr int OFF_SC = HDR;                                  // scale[64], shift[64]
  static constexpr int OFF_W = OFF_SC + 2 * 64 * 4;
  static constexpr int OFF_A = OFF_W + 2 * W_PLANE;                   // two operand buffers
  static constexpr int OFF_AUD = OFF_A + 2 * A_BUF;                   // FUSE1: two audio spans + conv1 weights
  static constexpr int OFF_W1 = OFF_AUD + (FUSE1 ? 2 * AUD_FLOATS * 4 : 0);
  static constexpr int OFF_RED = OFF_W1 + (FUSE1 ? KT * C1 * 4 : 0);
  static constexpr int SMEM = OFF_RED + 2 * COUT * 8;
  static constexpr uint32_t NCOLS = 2 * COUT <= 32 ? 32 : 2 * COUT <= 64 ? 64 : 2 * COUT <= 128 ? 128 : 256;
  static_assert(CIN % 16 == 0 && COUT % 16 == 0 && COUT <= 128 && CIN <= 64, "unsupported channel counts");
  static_assert(!FUSE1 || CIN == C1, "the fused first layer produces 16 channels");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// header: mbarriers (8 bytes each) at these byte offsets
constexpr int BAR_FULL = 0, BAR_EMPTY = 16, BAR_TFULL = 32, BAR_TEMPTY = 48, BAR_W = 64, TMEM_SLOT = 80;

template <int CIN, int COUT, bool FUSE1>
__global__ void __launch_bounds__(THREADS, 1) wav_conv_kernel(Params p) {
  using G = Geo<CIN, COUT, FUSE1>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t sbase = smem_u32(smem);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + TMEM_SLOT);
  float* sc = reinterpret_cast<float*>(smem + G::OFF_SC);
  float* sh = sc + 64;
  unsigned char* a_base = smem + G::OFF_A;
  float* aud = reinterpret_cast<float*>(smem + G::OFF_AUD);
  float* w1s = reinterpret_cast<float*>(smem + G::OFF_W1);     // [KT][16]
  double* red = reinterpret_cast<double*>(smem + G::OFF_RED);  // [2][COUT]

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(sbase + BAR_FULL + 8 * b, FUSE1 ? NPW : NPW / 2);   // one elected arrival per staging warp
      mbar_init(sbase + BAR_EMPTY + 8 * b, 1);    // tcgen05.commit
      mbar_init(sbase + BAR_TFULL + 8 * b, 1);    // tcgen05.commit
      mbar_init(sbase + BAR_TEMPTY + 8 * b, 8);   // one elected arrival per epilogue warp
    }
    mbar_init(sbase + BAR_W, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + TMEM_SLOT, G::NCOLS);
  // ---- prologue: BatchNorm of the input as scale / shift (CTA 0 updates the running statistics once); the fused first
  //      layer folds conv1's bias into the shift
  if (tid < CIN) {
    float a, b;
    bn_scale_shift(p.bn, tid, CIN, blockIdx.x == 0, a, b);
    if (FUSE1) b = fmaf(__ldg(p.b1 + tid), a, b);
    sc[tid] = a; sh[tid] = b;
  }
  for (int i = tid; i < 2 * COUT; i += THREADS) red[i] = 0.0;
  if (FUSE1)
    for (int i = tid; i < KT * C1; i += THREADS) w1s[i] = __ldg(p.w1 + (i % C1) * KT + i / C1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = p.tiles > (int)blockIdx.x ? (p.tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const float slope = p.bn.slope;   // 0 <= slope <= 1: LeakyReLU(v) = max(v, slope * v)

  if (warp_u < NPW) {
    // =================================================================================================== producers
    if (FUSE1) {
      // all staging warps work on the same tile: conv1 (fp32 SIMT, FFMA2) + BN1 + LeakyReLU -> operand images
      auto audio_prefetch = [&](int tile, int abuf) {
        // audio span of the tile's rows: at most two clips (pitch > R); second clip's span starts at seg1
        const int a0 = tile * TM;
        const int c0 = a0 / p.pitch, j0 = a0 - c0 * p.pitch;
        const int n0 = min(R, p.pitch - j0);
        const uint32_t dst = smem_u32(aud + abuf * AUD_FLOATS);
        const float* src0 = p.x + (long)c0 * p.L;
        for (int i = tid; i < S1 * S * n0 + KT; i += NPROD) {
          const int s = S1 * S * j0 - PAD1 + i;
          const bool ok = s >= 0 && s < p.L;
          cp_async4(dst + 4u * i, src0 + (ok ? s : 0), ok);
        }
        if (n0 < R) {
          const int n1 = R - n0, seg1 = S1 * S * n0 + 16;
          const bool clip_ok = c0 + 1 < p.B;
          const float* src1 = p.x + (long)(clip_ok ? c0 + 1 : c0) * p.L;
          for (int i = tid; i < S1 * S * n1 + KT; i += NPROD) {
            const int s = i - PAD1;
            const bool ok = clip_ok && s >= 0 && s < p.L;
            cp_async4(dst + 4u * (seg1 + i), src1 + (ok ? s : 0), ok);
          }
        }
      };
      if (my_tiles > 0) audio_prefetch(blockIdx.x, 0);
      cp_async_commit();
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int tile = blockIdx.x + tl * gridDim.x;
        const int a0 = tile * TM;
        cp_async_wait<0>();
        asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");   // every warp's copies landed; conv1 of tile tl-1 is done
        if (tl + 1 < my_tiles) audio_prefetch(tile + gridDim.x, (tl + 1) & 1);
        cp_async_commit();
        const int buf = tl & 1;
        if (tl >= 2) mbar_wait(sbase + BAR_EMPTY + 8 * buf, ((tl >> 1) - 1) & 1);
        unsigned char* a_hi = a_base + buf * A_BUF;
        unsigned char* a_lo = a_hi + A_PLANE;
        const float* au = aud + (tl & 1) * AUD_FLOATS;
        const int c0 = a0 / p.pitch, j0 = a0 - c0 * p.pitch;
        const int n0 = min(R, p.pitch - j0);
        const int seg1 = S1 * S * n0 + 16;
        // thread = local pixels tid and tid + NPROD (consecutive threads, consecutive pixels: sample stride 5,
        // conflict-free shared loads); local pixel lp = 6 * row + phase
        int off[2], dst[2]; bool ok[2], real[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int lp = tid + NPROD * i;
          ok[i] = lp < R * S;
          const int row = lp / S, r = lp - row * S;
          const bool second = row >= n0;
          const int pix = second ? lp - S * n0 : S * j0 + lp;
          real[i] = ok[i] && (c0 + (second ? 1 : 0)) < p.B && pix < p.Lin;
          off[i] = real[i] ? (second ? seg1 + S1 * pix : S1 * lp) : 0;
          dst[i] = (r * PST + row) * 16;
        }
        float2 acc[2][C1 / 2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int c = 0; c < C1 / 2; ++c) acc[i][c] = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < KT; ++t) {
          float2 wv[C1 / 2];
#pragma unroll
          for (int c4 = 0; c4 < C1 / 4; ++c4) {
            const float4 q = *reinterpret_cast<const float4*>(w1s + t * C1 + 4 * c4);
            wv[2 * c4] = make_float2(q.x, q.y); wv[2 * c4 + 1] = make_float2(q.z, q.w);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float xv = au[off[i] + t];
            const float2 x2 = make_float2(xv, xv);
#pragma unroll
            for (int c = 0; c < C1 / 2; ++c) acc[i][c] = __ffma2_rn(x2, wv[c], acc[i][c]);
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (!ok[i]) continue;
#pragma unroll
          for (int kc = 0; kc < 2; ++kc) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = kc * 8 + e;
              const float a = (e & 1) ? acc[i][c >> 1].y : acc[i][c >> 1].x;
              const float tv = fmaf(a, sc[c], sh[c]);
              v[e] = real[i] ? fmaxf(tv, tv * slope) : 0.f;
            }
            uint4 hi, lo;
            pack8w(v, hi, lo);
            *reinterpret_cast<uint4*>(a_hi + dst[i] + kc * RC * 16) = hi;
            if (p.x3) *reinterpret_cast<uint4*>(a_lo + dst[i] + kc * RC * 16) = lo;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(sbase + BAR_FULL + 8 * buf);
      }
      cp_async_wait<0>();
    } else {
      // two staging groups of NPW/2 warps: group g stages the items n = g (mod 2) into buffer g, so the loads of one
      // item are in flight while the other group converts (raw producer output -> BN + LeakyReLU -> operand images).
      // item = (local pixel, 8-channel chunk): consecutive threads read consecutive 32-byte segments
      constexpr int GT = (NPW / 2) * 32;                  // threads per group
      constexpr int ITEMS = R * S * 2;
      constexpr int PER = (ITEMS + GT - 1) / GT;          // items per thread (7)
      const int g = warp_u >= NPW / 2 ? 1 : 0, gtid = tid - g * GT;
      unsigned char* a_hi = a_base + g * A_BUF;
      unsigned char* a_lo = a_hi + A_PLANE;
      const int n_items = my_tiles * G::NPASS;
      int use = 0;
      for (int n = g; n < n_items; n += 2, ++use) {
        const int tl = n / G::NPASS, pass = n - tl * G::NPASS;
        const int a0 = (blockIdx.x + tl * gridDim.x) * TM;
        const int ch0 = pass * 16;
        if (use >= 1) mbar_wait(sbase + BAR_EMPTY + 8 * g, (use - 1) & 1);
#pragma unroll
        for (int b0 = 0; b0 < PER; b0 += 4) {
          float4 va[4], vb[4]; int dst[4]; bool ok[4], real[4]; int chv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int it = gtid + GT * (b0 + i);
            ok[i] = (b0 + i < PER) && it < ITEMS;
            const int kc = it & 1, lp = it >> 1;
            const int row = lp / S, r = lp - row * S;
            const int gr = a0 + row;
            const int c = gr / p.pitch, j = gr - c * p.pitch;
            const int pix = S * j + r;
            real[i] = ok[i] && c < p.B && pix < p.Lin;
            chv[i] = ch0 + kc * 8;
            dst[i] = (r * PST + kc * RC + row) * 16;
            if (real[i]) {
              const float4* src = reinterpret_cast<const float4*>(p.x + ((long)c * p.Lin + pix) * CIN + chv[i]);
              va[i] = __ldg(src); vb[i] = __ldg(src + 1);
            } else {
              va[i] = make_float4(0.f, 0.f, 0.f, 0.f); vb[i] = va[i];
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!ok[i]) continue;
            float v[8] = {va[i].x, va[i].y, va[i].z, va[i].w, vb[i].x, vb[i].y, vb[i].z, vb[i].w};
            if (real[i]) {
              const float4* s4 = reinterpret_cast<const float4*>(sc + chv[i]);
              const float4* h4 = reinterpret_cast<const float4*>(sh + chv[i]);
              const float4 s0 = s4[0], s1v = s4[1], h0 = h4[0], h1 = h4[1];
              const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1v.x, s1v.y, s1v.z, s1v.w};
              const float shv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float tv = fmaf(v[e], scv[e], shv[e]);
                v[e] = fmaxf(tv, tv * slope);
              }
            }
            uint4 hi, lo;
            pack8w(v, hi, lo);
            *reinterpret_cast<uint4*>(a_hi + dst[i]) = hi;
            if (p.x3) *reinterpret_cast<uint4*>(a_lo + dst[i]) = lo;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(sbase + BAR_FULL + 8 * g);
      }
    }
  } else if (warp_u == NPW) {
    // ====================================================================================================== issuer
    if (elect_one()) {
      // the layer's whole weight image: one expect_tx, bulk copies of <= 32 KB
      constexpr uint32_t WB = 2u * G::W_PLANE;
      mbar_expect_tx(sbase + BAR_W, WB);
      for (uint32_t o = 0; o < WB; o += 32768u)
        bulk_g2s(sbase + G::OFF_W + o, p.wpk + o, WB - o < 32768u ? WB - o : 32768u, sbase + BAR_W);
      mbar_wait(sbase + BAR_W, 0);
      const uint32_t idesc = make_idesc(COUT);
      const uint32_t sw = sbase + G::OFF_W;
      int n = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int acc = tl & 1;
        const uint32_t d = tmem_base + (uint32_t)(acc * COUT);
        if (tl >= 2) mbar_wait(sbase + BAR_TEMPTY + 8 * acc, ((tl >> 1) - 1) & 1);
        for (int pass = 0; pass < G::NPASS; ++pass, ++n) {
          const int buf = n & 1;
          mbar_wait(sbase + BAR_FULL + 8 * buf, (n >> 1) & 1);
          tc_fence_after();
          const uint32_t sa = sbase + G::OFF_A + buf * A_BUF;
          uint32_t cnt = pass;
#pragma unroll 1
          for (int t = 0; t < KT; ++t) {
            const int q = t / S, r = t - q * S;
            const uint32_t ah = sa + (uint32_t)((r * PST + q) * 16), al = ah + (uint32_t)A_PLANE;
            const uint32_t wh = sw + (uint32_t)(((t * G::KCT + 2 * pass) * COUT) * 16), wl = wh + (uint32_t)G::W_PLANE;
            const uint64_t dah = make_desc(ah, RC * 16, 128), dwh = make_desc(wh, COUT * 16, 128);
            if (p.x3) {
              mma_bf16(d, make_desc(al, RC * 16, 128), dwh, idesc, cnt ? 1u : 0u); ++cnt;
              mma_bf16(d, dah, make_desc(wl, COUT * 16, 128), idesc, 1u); ++cnt;
            }
            mma_bf16(d, dah, dwh, idesc, cnt ? 1u : 0u); ++cnt;
          }
          mma_commit(sbase + BAR_EMPTY + 8 * buf);
        }
        mma_commit(sbase + BAR_TFULL + 8 * acc);
      }
    }
  } else {
    // ==================================================================================================== epilogue
    // 8 warps: TMEM lane quadrant = warp & 3 (hardware rule), channel half = (warp - NPW - 1) >> 2
    constexpr int CH = COUT / 2;
    float s1[CH], s2[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) s1[c] = s2[c] = 0.f;
    const int quad = warp & 3, half = (warp - NPW - 1) >> 2;
    const int c_beg = half * CH;
    const bool vec_dst = (p.ldy & 7) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int tile = blockIdx.x + tl * gridDim.x;
      const int acc = tl & 1;
      mbar_wait(sbase + BAR_TFULL + 8 * acc, (tl >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * COUT + c_beg);
      const int a = tile * TM + quad * 32 + lane;
      const int c = a / p.pitch, o = a - c * p.pitch;
      const bool ok = c < p.B && o < p.Lo;
      float* dst = p.y + ((long)c * p.Lo + o) * p.ldy + c_beg;
#pragma unroll
      for (int cg = 0; cg < CH / 8; ++cg) {
        float t8[8];
        tmem_ld8w(t_addr + cg * 8, t8);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cg == CH / 8 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(sbase + BAR_TEMPTY + 8 * acc);   // the accumulator may be overwritten
        }
        if (ok) {
          float val[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float d = t8[i];
            s1[cg * 8 + i] += d; s2[cg * 8 + i] = fmaf(d, d, s2[cg * 8 + i]);
            val[i] = d + __ldg(p.bias + c_beg + cg * 8 + i);
          }
          if (vec_dst) {
            asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"f"(val[0]), "f"(val[1]), "f"(val[2]),
                         "f"(val[3]), "f"(val[4]), "f"(val[5]), "f"(val[6]), "f"(val[7]), "l"(dst + cg * 8)
                         : "memory");
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[cg * 8 + i] = val[i];
          }
        }
      }
    }
    // per-channel sums of this layer's raw output (relative to its bias) for the next BatchNorm
    if (p.sums_out) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float a = s2ag_warp_sum(s1[c]), b = s2ag_warp_sum(s2[c]);
        if (lane == 0) { atomicAdd(&red[c_beg + c], (double)a); atomicAdd(&red[COUT + c_beg + c], (double)b); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.sums_out)
    for (int i = tid; i < 2 * COUT; i += THREADS) atomicAdd(p.sums_out + i, red[i]);
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
  p.total_rows = p.B * p.pitch;
  p.tiles = (p.total_rows + TM - 1) / TM;
  p.x3 = umma::g_precision == 0 ? 1 : 0;
  int grid = sms < p.tiles ? sms : p.tiles;
  S2AG_LAUNCH(kfn, grid, THREADS, G::SMEM, stream, p);
  return 0;
}

constexpr long WPK2 = 2L * KT * 2 * 32 * 16, WPK3 = 2L * KT * 4 * 64 * 16, WPK4 = 2L * KT * 8 * 32 * 16;

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
  if (L3 < KT) return 0;
  return (long)B * ((long)L2 * 32 + (long)L3 * 64) + 2 * (2 * 16 + 2 * 32 + 2 * 64) + (WPK2 + WPK3 + WPK4) / 4 + 64;
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
  S2AG_CHECK_ARG(ldy >= 32 && (long)B * L1 < (1L << 31) && (long)B * L < (1L << 40));
  S2AG_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 31) == 0);
  S2AG_CHECK_ARG(slope >= 0.f && slope <= 1.f);   // LeakyReLU evaluated as max(v, slope * v)
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = s2ag_sm_count();
  // workspace: raw conv2 output, raw conv3 output, sums (doubles) of conv1 / conv2 / conv3, packed weight images
  float* y2 = ws;
  float* y3 = y2 + (long)B * L2 * 32;
  double* sums = reinterpret_cast<double*>(y3 + (long)B * L3 * 64);
  double* sums1 = sums; double* sums2 = sums + 2 * 16; double* sums3 = sums2 + 2 * 32;
  unsigned char* wpk2 = reinterpret_cast<unsigned char*>(sums + 2 * (16 + 32 + 64));
  unsigned char* wpk3 = wpk2 + WPK2;
  unsigned char* wpk4 = wpk3 + WPK3;
  S2AG_CHECK_ARG((reinterpret_cast<uintptr_t>(wpk2) & 15) == 0);
  {
    if (training) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (16 + 32 + 64), st);
    PrepParams pp;
    pp.audio = audio; pp.B = B; pp.L = L; pp.L1 = L1; pp.w1 = conv_w[0]; pp.sums = sums1;
    pp.items_per_clip = (L1 + K1_PIX - 1) / K1_PIX;
    int sb = training ? B * pp.items_per_clip : 0;
    if (sb > 2 * sms) sb = 2 * sms;
    pp.stat_blocks = sb;
    pp.job[0] = {conv_w[1], wpk2, 16, 32}; pp.job[1] = {conv_w[2], wpk3, 32, 64}; pp.job[2] = {conv_w[3], wpk4, 64, 32};
    auto k1 = &wav_prep_kernel;
    S2AG_LAUNCH(k1, sb + 8, 256, 0, stream, pp);
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
  p.w1 = conv_w[0]; p.b1 = conv_b[0]; p.wpk = wpk2; p.bias = conv_b[1]; p.y = y2; p.ldy = 32;
  p.sums_out = training ? sums2 : nullptr; p.bn = bn_in(0, sums1, (long)B * L1);
  S2AG_CHECK_ARG(p.pitch > R);
  if (launch_layer<16, 32, true>(p, sms, stream)) { s2ag_set_error("wavencoder: shared memory attribute"); return S2AG_ERR_LAUNCH; }
  // BN2 -> conv3
  p.x = y2; p.Lin = L2; p.Lo = L3; p.pitch = (L2 + S - 1) / S; p.w1 = nullptr; p.b1 = nullptr;
  p.wpk = wpk3; p.bias = conv_b[2]; p.y = y3; p.ldy = 64; p.sums_out = training ? sums3 : nullptr;
  p.bn = bn_in(1, sums2, (long)B * L2);
  if (launch_layer<32, 64, false>(p, sms, stream)) { s2ag_set_error("wavencoder: shared memory attribute"); return S2AG_ERR_LAUNCH; }
  // BN3 -> conv4 -> features
  p.x = y3; p.Lin = L3; p.Lo = L4; p.pitch = (L3 + S - 1) / S; p.wpk = wpk4; p.bias = conv_b[3]; p.y = y; p.ldy = ldy;
  p.sums_out = nullptr; p.bn = bn_in(2, sums3, (long)B * L3);
  if (launch_layer<64, 32, false>(p, sms, stream)) { s2ag_set_error("wavencoder: shared memory attribute"); return S2AG_ERR_LAUNCH; }
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
