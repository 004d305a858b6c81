// Causal-TCN residual block forward (net/tcn.py:16-46: weight_norm(Conv1d) Chomp1d ReLU Dropout, twice, + residual, ReLU;
// kernel 2, n_in == n_out -- the configuration net/multimodal_context_net_v2.py:75-76 instantiates) as ONE tcgen05 kernel
// for sm_100a, plus a packing kernel that folds weight_norm into the tensor-core operand images of both convolutions.
//
// The two-launch form (tcn.cu) wrote y1 = drop(relu(conv1(x))) to HBM, read it back as the gathered two-tap operand of
// the second GEMM and paid a prologue / epilogue / operand conversion per launch (6 launches per block with weight_norm
// and packing).  Here a CTA owns whole clips, so the causal halo is local and y1 never leaves the SM:
//   * rows of a tile = G clips laid out with a pitch of T + d rows (d zero rows after each clip: the causal padding of
//     the next one) under LEAD = d leading zero rows; the activation image [8-channel chunk][row][16 B] (bf16 hi/lo,
//     K-major UMMA layout) is staged ONCE and the tap "t - d" is the same image with the descriptor start d rows earlier
//     (shifted-window contraction, as in umma_conv.cu);
//   * both weight tensors stream through a 5-slot TMA ring (cp.async.bulk) of pre-packed [tap][16 channels] k-step images
//     of one column half (2 planes x 2 chunks x 160 x 16 B = 10 KB per slot at C = 300).  MEASURED LIMIT (tools/
//     tcn_timeline.cu): 50 KB of ring / ~2.5 k cycles of bulk-copy round trip = 20 B/clk per SM, i.e. 19 k cycles per
//     column half against 8.3 k cycles of MMA issue -- the kernel is bound by the weight stream (every tile of 3 clips
//     pulls the whole 1.5 MB operand image through its SM); a second ingest path (worker-warp LDG) and a per-CTA
//     rotated k order were tried and change nothing;
//   * conv1 accumulates in TMEM (Cpad fp32 columns as two column halves, N <= 256 each, one after the other; the weight
//     stream is ordered [half][k-step]); the epilogue applies bias, ReLU, dropout and writes y1 as the bf16 hi/lo operand
//     image of conv2 over the x image (and to HBM only when the backward pass needs it); conv2 accumulates into the
//     same TMEM columns; its epilogue adds bias, ReLU, dropout, the residual x (re-read from L2) and the final ReLU --
//     for column half 0 while the MMAs of half 1 are still running.
// Roles: warps 0-15 stage x and run both epilogues (4 per TMEM lane quadrant, a quarter of the columns each), warps 16
// and 17 issue the MMAs (alternate ring items), warp 18 drives the TMA ring.
#include "s2ag.h"
#include "gemm_umma.cuh"
#include "gemm_umma_packed.cuh"

namespace s2ag {
namespace tcnf {

using namespace s2ag::umma;

constexpr int NISSUE = 3;              // MMA issuer threads of the single-CTA kernel (one warp each)
constexpr int TM = 128, NWW = 16, NWORK = NWW * 32, THREADS = NWORK + 32 * (NISSUE + 1), HDR = 1024;   // 16 worker warps + issuers + TMA
constexpr int NSTAGE_MAX = 12, NSTAGE_2 = 8;   // weight-ring slots: single CTA: as many 10 KB slots as fit (5..12) / CTA pair: 5 KB each, one relay warp per slot
// mbarriers (8 bytes each): weights landed (own copy) | ring slot free | weights landed in the peer CTA (pair leader only) |
// accumulator column half complete | operand image staged
constexpr int BAR_WFULL = 0, BAR_WEMPTY = 128, BAR_PFULL = 256, BAR_ACC = 384, BAR_AREADY = 400, BAR_ORD = 408, TMEM_SLOT = 416;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// long waits of many warps (the 16 worker warps while a convolution's MMAs run): poll with a sleep in between, so that
// the polling does not compete with the tensor core's operand reads for the shared-memory pipe
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(200);
    if (++spins > (1u << 22)) __trap();
  }
}
// CTA-pair forms (semantics checked by tools/mma2_probe.cu): M = 256 over the pair, each CTA supplies its 128 rows of A and
// its half of B's rows from the SAME shared-memory offsets; the commit arrives on the barrier at that offset in both CTAs
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_pair(int n) {   // M = 256
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

struct Params {
  const float* x; const float* b1; const float* b2;
  const unsigned char* wpk;     // [conv][column half h][k-step = tap * KS + s][plane hi|lo][2 chunks][N_h][16 B]
  float* y1; float* y2; float* out;   // y1 / y2 may be NULL (no backward pass)
  int B, T, C, d, G, pitch, tiles;
  int Kc, Cpad, KS, N0, N1, Rv, RCH, nstage; // chunks of 8 channels, padded channels, k-steps per tap, MMA column split,
                                             // anchor rows that can be valid (G * pitch), image rows per chunk, ring slots
  float p_drop; unsigned long long seed; const unsigned long long* seed_dev;
  int x3;
};

__device__ __forceinline__ void pack8t(const float (&v)[8], uint4& hi, uint4& lo) {
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
__device__ __forceinline__ void tmem_ld8t(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}

// weight_norm folded into the operand packing: block = (output channel n, conv); w[n][j][c] = v[n][c][j] * g[n] / |v[n]|
// written (a) as fp32 tap-major [n][j][c] + the norm (what the backward pass consumes) and (b) as bf16 hi/lo k-step
// images [tap j][s][plane][chunk 2][Cpad][16 B] (channels >= C and rows >= C zero).
__global__ void __launch_bounds__(128) tcn_pack_kernel(const float* __restrict__ v1, const float* __restrict__ g1,
                                                       const float* __restrict__ v2, const float* __restrict__ g2,
                                                       float* __restrict__ w1, float* __restrict__ w2,
                                                       float* __restrict__ n1, float* __restrict__ n2,
                                                       unsigned char* __restrict__ wpk, int C, int Cpad, int KS, int N0, int pair) {
  __shared__ float red[4];
  __shared__ float wrow[2 * 328];
  const int n = blockIdx.x, conv = blockIdx.y, tid = threadIdx.x;
  const float* v = conv ? v2 : v1; const float* g = conv ? g2 : g1;
  float* w = conv ? w2 : w1; float* nrm_out = conv ? n2 : n1;
  // column half of this row: h = 0 holds rows [0, N0), h = 1 rows [N0, Cpad); a stage of half h is 2 planes x 2 chunks x N_h
  const int hN = n < N0 ? 0 : 1, Nh = hN ? Cpad - N0 : N0, nn = n - hN * N0;
  const long stage_bytes = 2L * 2 * Nh * 16;
  unsigned char* img = wpk + (long)conv * 2 * KS * (2L * 2 * Cpad * 16) + (hN ? 2L * KS * (2L * 2 * N0 * 16) : 0);
  float sc = 0.f;
  if (n < C) {
    const float* vr = v + (long)n * C * 2;
    float s = 0.f;
    for (int i = tid; i < 2 * C; i += 128) { const float t = vr[i]; s = fmaf(t, t, s); }
    s = s2ag_warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    const float nrm = sqrtf(red[0] + red[1] + red[2] + red[3]);
    sc = g[n] / nrm;
    if (tid == 0) nrm_out[n] = nrm;
    for (int i = tid; i < 2 * C; i += 128) {
      const int c = i >> 1, j = i & 1;           // v is [c][j]
      const float val = vr[i] * sc;
      w[(long)n * 2 * C + j * C + c] = val;
      wrow[j * 328 + c] = val;
    }
  }
  for (int i = tid; i < 2 * 328; i += 128) { const int c = i % 328; if (n >= C || c >= C) wrow[i] = 0.f; }
  __syncthreads();
  // chunks of this row: (tap j, k-step s, chunk q) -> 8 channels 16 s + 8 q ..
  for (int it = tid; it < 2 * KS * 2; it += 128) {
    const int q = it & 1, s = (it >> 1) % KS, j = it / (2 * KS);
    float vals[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) vals[e] = wrow[j * 328 + s * 16 + q * 8 + e];
    uint4 hi, lo;
    pack8t(vals, hi, lo);
    unsigned char* dst;
    long plane;
    if (pair) {   // [CTA c][plane][chunk q][Nh/2 rows][16 B]: CTA c of a pair holds rows [c Nh/2, (c+1) Nh/2)
      const int hr = Nh / 2, c = nn / hr, r = nn - c * hr;
      plane = 2L * hr * 16;
      dst = img + (long)(j * KS + s) * stage_bytes + (long)c * (stage_bytes / 2) + ((long)q * hr + r) * 16;
    } else {
      plane = 2L * Nh * 16;
      dst = img + (long)(j * KS + s) * stage_bytes + ((long)q * Nh + nn) * 16;
    }
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + plane) = lo;
  }
}

#ifdef S2AG_TCN_TIMELINE
__device__ long long g_tcn_tl[3][16];   // [worker | issuer | producer][mark]: clock64 of CTA 0
#define TCN_MARK(role, slot) do { if (blockIdx.x == 0 && lane == 0) g_tcn_tl[role][slot] = clock64(); } while (0)
#else
#define TCN_MARK(role, slot) do { } while (0)
#endif

template <bool PAIR>
__global__ void __launch_bounds__(THREADS, 1) tcn_block_fused_kernel(Params p) {
  const int NSTAGE = PAIR ? NSTAGE_2 : p.nstage;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;       // PAIR: 2-CTA cluster, rank 0 issues the MMAs for both
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t sbase = smem_u32(smem);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + TMEM_SLOT);
  const int a_plane = p.Kc * p.RCH * 16;
  const int stage_slot = (2 * 2 * p.N0 * 16) / (PAIR ? 2 : 1);   // ring slot = this CTA's share of the larger column half's stage
  float* bias_s = reinterpret_cast<float*>(smem + HDR);            // b1[Cpad], b2[Cpad]
  unsigned char* a_hi = smem + HDR + 2 * p.Cpad * 4;
  unsigned char* a_lo = a_hi + a_plane;
  unsigned char* wst = a_lo + a_plane;                             // NSTAGE stages
  const int LEAD = p.d;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(sbase + BAR_WFULL + 8 * s, 1); mbar_init(sbase + BAR_WEMPTY + 8 * s, 1); mbar_init(sbase + BAR_PFULL + 8 * s, 1);
    }
    mbar_init(sbase + BAR_ACC, PAIR ? 1 : NISSUE);    // single-CTA kernel: every issuer thread commits
    mbar_init(sbase + BAR_ACC + 8, PAIR ? 1 : NISSUE);
    mbar_init(sbase + BAR_ORD, 1);
    mbar_init(sbase + BAR_AREADY, PAIR ? 2 * NWW : NWW);   // PAIR: the leader's barrier also collects the peer's workers
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + TMEM_SLOT), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(sbase + TMEM_SLOT, 512);
    }
  }
  for (int i = tid; i < 2 * p.Cpad; i += THREADS) {
    const int c = i % p.Cpad;
    bias_s[i] = c < p.C ? __ldg((i < p.Cpad ? p.b1 : p.b2) + c) : 0.f;
  }
  // rows that no epilogue ever writes (the LEAD zero rows above the first anchor) are cleared once
  for (int i = tid; i < p.Kc * LEAD; i += THREADS) {
    const int kc = i / LEAD, r = i - kc * LEAD;
    *reinterpret_cast<uint4*>(a_hi + (kc * p.RCH + r) * 16) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(a_lo + (kc * p.RCH + r) * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // the peer's barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PAIR: both CTAs of a pair run the same number of tiles (the odd one out is an empty tile: clips beyond the batch)
  const int units = PAIR ? (p.tiles + 1) / 2 : p.tiles, unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ustride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int my_tiles = units > unit0 ? (units - 1 - unit0) / ustride + 1 : 0;
  const uint32_t aready_bar = PAIR ? mapa(sbase + BAR_AREADY, 0u) : 0u;   // the leader's "images staged" barrier
  const int n_steps = 2 * p.KS;   // k-steps (stages) per convolution

  if (warp_u < NWW) {
    // ======================================================================================= workers / epilogues
    const unsigned long long seed1 = p.seed + ((p.p_drop > 0.f && p.seed_dev) ? p.seed_dev[0] : 0ull);
    const unsigned long long seed2 = seed1 + 0x1234567ull;
    const int quad = warp & 3, part = warp >> 2;      // TMEM lane quadrant (hardware rule: warp % 4), column quarter
    const int half = part >> 1;
    const int row = quad * 32 + lane;                 // anchor row of the epilogues (TMEM lane)
    const int gcl = row / p.pitch, t = row - gcl * p.pitch;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int tile = PAIR ? 2 * (unit0 + tl * ustride) + (int)rank : (int)blockIdx.x + tl * (int)gridDim.x;
      const int clip0 = tile * p.G;
      // ---- x -> operand image (item = (row, chunk), chunk fastest: coalesced 32-byte reads; RCH odd: conflict-free)
      const bool vec_x = (p.C & 3) == 0;
      if (warp == 0) TCN_MARK(0, 0);
      // only the G * pitch rows that can hold a clip (or its padding) are staged: the image is truncated there, the MMA
      // reads its 128 rows past the end into whatever follows (garbage rows of the accumulator, never stored)
      for (int base = 0; base < p.Rv * p.Kc; base += 4 * NWORK) {
        float4 va[4], vb[4]; int dst[4]; bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int it = base + tid + NWORK * i;
          ok[i] = it < p.Rv * p.Kc;
          const int kc = it % p.Kc, r = it / p.Kc;
          const int g = r / p.pitch, tt = r - g * p.pitch;
          dst[i] = (kc * p.RCH + LEAD + r) * 16;
          va[i] = make_float4(0.f, 0.f, 0.f, 0.f); vb[i] = va[i];
          if (ok[i] && g < p.G && tt < p.T && clip0 + g < p.B) {
            const float* src = p.x + ((long)(clip0 + g) * p.T + tt) * p.C + kc * 8;
            if (vec_x && kc * 8 + 8 <= p.C) {
              va[i] = __ldg(reinterpret_cast<const float4*>(src)); vb[i] = __ldg(reinterpret_cast<const float4*>(src) + 1);
            } else {
              float t8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) t8[e] = kc * 8 + e < p.C ? __ldg(src + e) : 0.f;
              va[i] = make_float4(t8[0], t8[1], t8[2], t8[3]); vb[i] = make_float4(t8[4], t8[5], t8[6], t8[7]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!ok[i]) continue;
          const float v[8] = {va[i].x, va[i].y, va[i].z, va[i].w, vb[i].x, vb[i].y, vb[i].z, vb[i].w};
          uint4 hi, lo;
          pack8t(v, hi, lo);
          *reinterpret_cast<uint4*>(a_hi + dst[i]) = hi;
          if (p.x3) *reinterpret_cast<uint4*>(a_lo + dst[i]) = lo;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_remote(aready_bar); else mbar_arrive_cta(sbase + BAR_AREADY); }
      const bool valid = gcl < p.G && t < p.T && clip0 + gcl < p.B;
      const long grow = ((long)(clip0 + gcl) * p.T + t) * p.C;     // global element offset of this row
      const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
      const int nch = p.Kc;                                         // chunks of 8 output channels
      // column halves follow the accumulator split (N0 | N1) so that the epilogue of half 0 overlaps the MMAs of half 1
      const int split = p.N1 > 0 ? p.N0 / 8 : (nch + 1) / 2;
      const int h_beg = half ? split : 0, h_end = half ? nch : split;          // this half's chunks, then its two quarters
      const int h_mid = h_beg + (h_end - h_beg + 1) / 2;
      const int ch_beg = (part & 1) ? h_mid : h_beg, ch_end = (part & 1) ? h_end : h_mid;
      // peer CTA of a pair: the leader's issuer must know that THIS CTA's share of a ring item has landed.  Worker warps
      // 8..15 (column half 1: they wait for the end of either convolution anyway) each relay one ring slot: wait for the
      // local "landed" barrier, arrive on the leader's PFULL barrier of that slot.  (One thread relaying all slots costs
      // ~850 cycles per item, lanes of one warp polling different barriers serialise on mbarrier.try_wait.)
      auto relay = [&](int conv) {
        if (!PAIR || rank != 1 || warp < NWW / 2) return;
        const int slot = warp - NWW / 2;
        const int per_conv = (p.N1 > 0 ? 2 : 1) * n_steps;
        const int n_beg = (tl * 2 + conv) * per_conv, n_end = n_beg + per_conv;
        if (lane == 0) {
          const uint32_t remote = mapa(sbase + BAR_PFULL + 8 * slot, 0u);
          int n = n_beg + ((slot - n_beg) % NSTAGE + NSTAGE) % NSTAGE;      // first item of this conv in my slot
          for (; n < n_end; n += NSTAGE) {
            mbar_wait(sbase + BAR_WFULL + 8 * slot, (n / NSTAGE) & 1);
            mbar_arrive_remote(remote);
          }
        }
        __syncwarp();
      };
      // conv1's epilogue overwrites the x image that BOTH column halves of conv1 read: it waits for the last half.
      // conv2's epilogue only writes global memory: half 0 starts as soon as its columns are complete.
      const uint32_t acc_bar1 = sbase + BAR_ACC + (p.N1 > 0 ? 8 : 0);
      const uint32_t acc_bar2 = sbase + BAR_ACC + ((p.N1 > 0 && half) ? 8 : 0);
      // ---- epilogue 1: y1 = drop(relu(acc + b1)) -> operand image of conv2 (zeros in the padding rows)
      if (warp == 0) TCN_MARK(0, 1);
      relay(0);
      mbar_wait_sleep(acc_bar1, 0u);
      if (warp == 0) TCN_MARK(0, 2);
      tc_fence_after();
      for (int ch = ch_beg; ch < ch_end; ++ch) {
        float v[8];
        tmem_ld8t(t_lane + (uint32_t)(ch * 8), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = ch * 8 + e;
          float a = v[e] + bias_s[c];
          a = a > 0.f ? a : 0.f;
          if (p.p_drop > 0.f) a *= s2ag_dropout_scale(seed1, (unsigned long long)(grow + c), p.p_drop);
          v[e] = (valid && c < p.C) ? a : 0.f;
        }
        if (p.y1 && valid) {
          if (ch * 8 + 8 <= p.C && (p.C & 3) == 0) {
            *reinterpret_cast<float4*>(p.y1 + grow + ch * 8) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(p.y1 + grow + ch * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (ch * 8 + e < p.C) p.y1[grow + ch * 8 + e] = v[e];
          }
        }
        if (row < p.Rv) {
          uint4 hi, lo;
          pack8t(v, hi, lo);
          const int dst = (ch * p.RCH + LEAD + row) * 16;
          *reinterpret_cast<uint4*>(a_hi + dst) = hi;
          if (p.x3) *reinterpret_cast<uint4*>(a_lo + dst) = lo;
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_remote(aready_bar); else mbar_arrive_cta(sbase + BAR_AREADY); }
      // ---- epilogue 2: y2 = drop(relu(acc + b2)); out = relu(y2 + x)
      if (warp == 0) TCN_MARK(0, 3);
      relay(1);
      mbar_wait_sleep(acc_bar2, 1u);
      if (warp == 0) TCN_MARK(0, 4);
      tc_fence_after();
      auto load_res = [&](int ch, float4& ra, float4& rb) {
        ra = make_float4(0.f, 0.f, 0.f, 0.f); rb = ra;
        if (!valid || ch >= ch_end) return;
        const float* src = p.x + grow + ch * 8;
        if (vec_x && ch * 8 + 8 <= p.C) {
          ra = __ldg(reinterpret_cast<const float4*>(src)); rb = __ldg(reinterpret_cast<const float4*>(src) + 1);
        } else {
          float t8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t8[e] = ch * 8 + e < p.C ? __ldg(src + e) : 0.f;
          ra = make_float4(t8[0], t8[1], t8[2], t8[3]); rb = make_float4(t8[4], t8[5], t8[6], t8[7]);
        }
      };
      float4 ra, rb;
      load_res(ch_beg, ra, rb);
      for (int ch = ch_beg; ch < ch_end; ++ch) {
        float v[8];
        const float res[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
        load_res(ch + 1, ra, rb);          // the next chunk's residual is in flight during this chunk's TMEM read
        tmem_ld8t(t_lane + (uint32_t)(ch * 8), v);
        if (!valid) continue;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = ch * 8 + e;
          float a = v[e] + bias_s[p.Cpad + c];
          a = a > 0.f ? a : 0.f;
          if (p.p_drop > 0.f) a *= s2ag_dropout_scale(seed2, (unsigned long long)(grow + c), p.p_drop);
          v[e] = a;
          const float s = a + res[e];
          o[e] = s > 0.f ? s : 0.f;
        }
        if (ch * 8 + 8 <= p.C && vec_x) {
          if (p.y2) {
            *reinterpret_cast<float4*>(p.y2 + grow + ch * 8) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(p.y2 + grow + ch * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          *reinterpret_cast<float4*>(p.out + grow + ch * 8) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(p.out + grow + ch * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (ch * 8 + e < p.C) {
              if (p.y2) p.y2[grow + ch * 8 + e] = v[e];
              p.out[grow + ch * 8 + e] = o[e];
            }
        }
      }
      if (warp == 0) TCN_MARK(0, 5);
      tc_fence_before();
      // the image and the accumulator are reused by the next tile: all epilogue warps must be done reading
      asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory");
    }
  } else if (warp_u >= NWW && warp_u < NWW + NISSUE) {
    // ===================================================================================================== issuers
    // Single-CTA kernel: TWO issuer threads take the ring items alternately.  Measured (tools/tcn_timeline.cu): one thread
    // needs ~470 cycles per item -- a tcgen05.mma issue blocks until the pipe accepts it (~73 cycles each at N = 160) and
    // the mbarrier wait for the next item (~180 cycles) starts only after the third issue -- against 220 cycles of MMA
    // time; with two threads one thread's wait overlaps the other's issues.  Both accumulate into the same TMEM
    // columns: the thread that owns the first k-step of a column half (accumulate = 0) releases the other through the
    // ORD barrier once that MMA is issued (the pipe executes in issue order).  PAIR: one issuer (the cluster leader).
    const int iss = warp_u - NWW;
    constexpr int NISS = PAIR ? 1 : NISSUE;
    if (iss < NISS && elect_one()) {
      if (!PAIR || rank == 0) {
        const uint32_t idesc0 = PAIR ? make_idesc_pair(p.N0) : make_idesc(p.N0);
        const uint32_t idesc1 = PAIR ? make_idesc_pair(p.N1 > 0 ? p.N1 : 16) : make_idesc(p.N1 > 0 ? p.N1 : 16);
        uint32_t aphase = 0, ophase = 0;
        uint32_t n = 0;                                // global ring item counter (both issuers count all items)
        int stage = 0; uint32_t sphase = 0;            // ring slot / phase parity of item n, advanced incrementally
        for (int tl = 0; tl < my_tiles; ++tl) {
          for (int conv = 0; conv < 2; ++conv) {
            if (iss == 0) TCN_MARK(1, conv * 4);
            mbar_wait(sbase + BAR_AREADY, aphase); aphase ^= 1u;
            if (iss == 0) TCN_MARK(1, conv * 4 + 1);
            tc_fence_after();
            for (int hN = 0; hN < 2; ++hN) {
              const int Nh = hN ? p.N1 : p.N0;
              if (Nh > 0) {
                const int Nl = PAIR ? Nh / 2 : Nh;             // B rows held by one CTA
                const uint32_t d = tmem_base + (hN ? (uint32_t)p.N0 : 0u);
                const uint32_t idesc = hN ? idesc1 : idesc0;
                const uint32_t plane = (uint32_t)(2 * Nl * 16);
                const bool first_mine = (int)(n % (uint32_t)NISS) == iss;      // who issues k-step 0 of this half
                if (NISS > 1 && !first_mine) { mbar_wait(sbase + BAR_ORD, ophase); }
                for (int ks = 0; ks < n_steps; ++ks, ++n) {
                  if ((int)(n % (uint32_t)NISS) == iss) {
                    mbar_wait(sbase + BAR_WFULL + 8 * stage, sphase);
                    if (PAIR) mbar_wait(sbase + BAR_PFULL + 8 * stage, sphase);   // the peer's share has landed too
                    tc_fence_after();
                    const int j = ks / p.KS, s = ks - j * p.KS;           // tap (0: t - d, 1: t), 16-channel group
                    const uint32_t ah = smem_u32(a_hi) + (uint32_t)((2 * s * p.RCH + LEAD - (j == 0 ? p.d : 0)) * 16);
                    const uint32_t al = ah + (uint32_t)a_plane;
                    const uint32_t wh = smem_u32(wst) + (uint32_t)(stage * stage_slot), wl = wh + plane;
                    const uint64_t dah = make_desc(ah, p.RCH * 16, 128), dwh = make_desc(wh, Nl * 16, 128);
                    const uint64_t dal = make_desc(al, p.RCH * 16, 128), dwl = make_desc(wl, Nl * 16, 128);
                    if (PAIR) {
                      if (p.x3) {
                        mma_bf16_pair(d, dal, dwh, idesc, ks ? 1u : 0u);
                        mma_bf16_pair(d, dah, dwl, idesc, 1u);
                        mma_bf16_pair(d, dah, dwh, idesc, 1u);
                      } else {
                        mma_bf16_pair(d, dah, dwh, idesc, ks ? 1u : 0u);
                      }
                      mma_commit_pair(sbase + BAR_WEMPTY + 8 * stage);
                    } else {
                      if (p.x3) {
                        mma_bf16(d, dal, dwh, idesc, ks ? 1u : 0u);
                        if (NISS > 1 && ks == 0) mbar_arrive_cta(sbase + BAR_ORD);   // the accumulator is initialised in issue order
                        mma_bf16(d, dah, dwl, idesc, 1u);
                        mma_bf16(d, dah, dwh, idesc, 1u);
                      } else {
                        mma_bf16(d, dah, dwh, idesc, ks ? 1u : 0u);
                        if (NISS > 1 && ks == 0) mbar_arrive_cta(sbase + BAR_ORD);
                      }
                      mma_commit(sbase + BAR_WEMPTY + 8 * stage);
                    }
                  }
                  if (++stage == NSTAGE) { stage = 0; sphase ^= 1u; }
                }
                ophase ^= 1u;
              }
              // this column half of the accumulator is complete (in both CTAs of a pair) once every issuer's MMAs are
              if (PAIR) mma_commit_pair(sbase + BAR_ACC + 8 * hN); else mma_commit(sbase + BAR_ACC + 8 * hN);
              if (iss == 0) TCN_MARK(1, conv * 4 + 2 + hN);
            }
          }
        }
      }
    }
  } else {
    // ================================================================================================ TMA producer
    if (elect_one()) {
      const long conv_bytes = (long)n_steps * (2L * 2 * p.Cpad * 16);
      int stage = 0; uint32_t ephase = 1; bool first_round = true;   // parity of the "slot free" phase to wait for
      for (int tl = 0; tl < my_tiles; ++tl) {
        for (int conv = 0; conv < 2; ++conv) {
          for (int hN = 0; hN < 2; ++hN) {
            const int Nh = hN ? p.N1 : p.N0;
            if (Nh == 0) continue;
            const uint32_t sb = (uint32_t)(2 * 2 * Nh * 16);          // stage bytes of the whole column half
            const uint32_t mine = PAIR ? sb / 2 : sb;                 // this CTA's share (PAIR: its half of the B rows)
            const unsigned char* base = p.wpk + conv * conv_bytes + (hN ? (long)n_steps * (2L * 2 * p.N0 * 16) : 0) +
                                        (PAIR ? (long)rank * mine : 0);
            for (int ks = 0; ks < n_steps; ++ks) {
              if (!first_round) mbar_wait(sbase + BAR_WEMPTY + 8 * stage, ephase);
              const unsigned char* src = base + (long)ks * sb;
              const uint32_t dst = smem_u32(wst) + (uint32_t)(stage * stage_slot);
              const uint32_t bar = sbase + BAR_WFULL + 8 * stage;
              // ONE bulk copy per item: the copy engine costs ~250-300 cycles per copy plus ~1 cycle per 28 bytes
              // (measured: two 5 KB copies per item 500 cycles, two 2.5 KB copies 640, one 19.5 KB copy 680)
              mbar_expect_tx(bar, mine);
              bulk_g2s(dst, src, mine, bar);
              if (++stage == NSTAGE) { stage = 0; ephase ^= 1u; first_round = false; }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) {
    cluster_sync_all();      // both CTAs are done with the pair's tensor memory and with each other's barriers
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  } else if (warp == 0) {
    tmem_dealloc(tmem_base, 512);
  }
}

static inline int rup(int a, int b) { return (a + b - 1) / b * b; }

int g_nstage_cap = 0;   // measurement aid (tools/tcn_timeline.cu): upper bound on the ring slots, 0 = as many as fit
struct Geom { int Kc, Cpad, KS, N0, N1, G, pitch, Rv, RCH, nstage; size_t smem; long wpk_bytes; bool ok; };
static Geom geom(int T, int C, int d) {
  Geom g;
  g.Cpad = rup(C, 16); g.Kc = g.Cpad / 8; g.KS = g.Cpad / 16;
  g.N0 = g.Cpad <= 256 ? g.Cpad : rup(g.Cpad / 2, 16);
  g.N1 = g.Cpad - g.N0;
  g.pitch = T + d; g.G = g.pitch <= TM ? TM / g.pitch : 0;
  g.Rv = g_nstage_cap < 0 ? TM : g.G * g.pitch; g.RCH = (g.Rv + d) | 1;          // image rows: d leading zero rows + the rows that can be valid
  const size_t fixed = HDR + 2 * (size_t)g.Cpad * 4 + 2 * (size_t)g.Kc * g.RCH * 16, slot = (size_t)2 * 2 * g.N0 * 16;
  // the MMA reads 128 rows from (at most) row d of the last chunk: the ring behind the image must cover that overhang
  const size_t overhang = g.RCH < TM + d ? (size_t)(TM + d - g.RCH) * 16 : 0;
  long ns = fixed < (size_t)227 * 1024 ? (long)(((size_t)227 * 1024 - fixed) / slot) : 0;
  if (ns > NSTAGE_MAX) ns = NSTAGE_MAX;
  if (g_nstage_cap > 0 && ns > g_nstage_cap) ns = g_nstage_cap;
  g.nstage = (int)ns;
  g.smem = fixed + (size_t)g.nstage * slot;
  g.wpk_bytes = 2L * 2 * g.KS * (2L * 2 * g.Cpad * 16);
  g.ok = g.G >= 1 && g.Cpad <= 512 && g.N1 <= 256 && g.nstage >= 3 && (size_t)g.nstage * slot >= overhang && C <= 320 && d >= 1;
  return g;
}

}  // namespace tcnf
}  // namespace s2ag

using namespace s2ag::tcnf;

// floats of workspace for s2ag_tcn_block_fused_fwd (the packed operand images of both convolutions); 0: unsupported shape
extern "C" long s2ag_tcn_fused_ws_floats(int T, int C, int dilation) {
  if (T <= 0 || C <= 0 || dilation <= 0) return 0;
  const Geom g = geom(T, C, dilation);
  return g.ok ? g.wpk_bytes / 4 + 8 : 0;
}

extern "C" int s2ag_tcn_block_fused_fwd(const float* x, const float* v1, const float* g1, const float* b1,
                                        const float* v2, const float* g2, const float* b2, float* w1, float* w2,
                                        float* n1, float* n2, float* y1, float* y2, float* out, float* ws, int B, int T,
                                        int C, int dilation, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                                        void* stream) {
  S2AG_CHECK_ARG(x && v1 && g1 && b1 && v2 && g2 && b2 && w1 && w2 && n1 && n2 && out && ws && B >= 0 && T > 0 && C > 0);
  S2AG_CHECK_ARG(dilation > 0 && p_drop >= 0.f && p_drop < 1.f && (reinterpret_cast<uintptr_t>(ws) & 15) == 0);
  const Geom g = geom(T, C, dilation);
  if (!g.ok) { s2ag_set_error("s2ag_tcn_block_fused_fwd: unsupported shape T=%d C=%d d=%d", T, C, dilation); return S2AG_ERR_UNSUPPORTED; }
  if (B == 0) return S2AG_OK;
  unsigned char* wpk = reinterpret_cast<unsigned char*>(ws);
  const int tiles = (B + g.G - 1) / g.G;
  // s2ag_debug_flags bit 8192: CTA pairs (tcgen05 cta_group::2, M = 256: each CTA streams only its half of the weight
  // rows).  Correct (parity-tested) but measured SLOWER than one CTA per tile (24 k vs 19 k cycles per column half): the
  // leader must learn that the peer's share of a ring item has landed, and that relay hop doubles the ring's round trip.
  const bool pair = tiles >= 2 && (s2ag::umma::g_dbg_flags & 8192);
  {
    auto kp = &tcn_pack_kernel;
    S2AG_LAUNCH(kp, dim3(g.Cpad, 2), 128, 0, stream, v1, g1, v2, g2, w1, w2, n1, n2, wpk, C, g.Cpad, g.KS, g.N0, pair ? 1 : 0);
  }
  const int sms = s2ag_sm_count();
  Params p;
  p.x = x; p.b1 = b1; p.b2 = b2; p.wpk = wpk; p.y1 = y1; p.y2 = y2; p.out = out;
  p.B = B; p.T = T; p.C = C; p.d = dilation; p.G = g.G; p.pitch = g.pitch; p.tiles = tiles;
  p.Kc = g.Kc; p.Cpad = g.Cpad; p.KS = g.KS; p.N0 = g.N0; p.N1 = g.N1; p.Rv = g.Rv; p.RCH = g.RCH; p.nstage = g.nstage;
  p.p_drop = p_drop; p.seed = seed; p.seed_dev = (const unsigned long long*)seed_dev;
  p.x3 = s2ag::umma::g_precision == 0 ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(&tcn_block_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(&tcn_block_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      s2ag_set_error("s2ag_tcn_block_fused_fwd: shared memory attribute"); return S2AG_ERR_LAUNCH;
    }
    attr_set = true;
  }
  if (pair) {
    const int units = (p.tiles + 1) / 2, max_pairs = sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (units < max_pairs ? units : max_pairs));
    cfg.blockDim = dim3(s2ag::tcnf::THREADS);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    ++g_s2ag_launches;
    if (cudaLaunchKernelEx(&cfg, &tcn_block_fused_kernel<true>, p) != cudaSuccess) {
      s2ag_set_error("s2ag_tcn_block_fused_fwd: cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      return S2AG_ERR_LAUNCH;
    }
  } else {
    auto kfn = &tcn_block_fused_kernel<false>;
    const int grid = p.tiles < sms ? p.tiles : sms;
    S2AG_LAUNCH(kfn, grid, s2ag::tcnf::THREADS, g.smem, stream, p);
  }
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}
