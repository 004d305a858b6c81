// Persistent, weight-stationary GRU recurrence for one bidirectional layer on tcgen05 (sm_100a).
// Replaces the T per-step launches of gru.cu for H <= 320 (nn.GRU call sites:
// net/multimodal_context_net_v2.py:480-481,541 (G), :281-282,333 (frozen tri-modal), :558-560,576 (D)).
//
// Decomposition.  grid = (S hidden slices of 16 units, batch tiles of 128 clips, 2 directions); every CTA
// stays resident for all T steps (<= 148 CTAs, one per SM: ~214 KB of shared memory each).
//   * W_hh stationary: the CTA's 48 rows of W_hh (gates r,z,n of its 16 hidden units) are split into bf16
//     hi/lo once and kept in shared memory in the canonical K-major UMMA layout [k-chunk][48][16 B].
//   * per step the CTA needs h_{t-1} of ALL hidden units of its 128 clips (the A operand,
//     [k-chunk][128][16 B] hi/lo): the slices publish their 16 new hidden units as a ready-made bf16 hi/lo
//     operand image in global memory (L2-resident, double-buffered by step parity); a per-(tile, direction)
//     arrival counter (release/acquire) orders producers and consumers; the image is copied L2 -> shared
//     memory with 16-byte ld.global.cg and consumed by tcgen05.mma (M=128, N=48, K=16; three MMAs per
//     k-step in the fp32-grade bf16x3 mode) into a 48-column TMEM accumulator.
//   * epilogue (thread = clip, 8 hidden units each): tcgen05.ld, gate math in fp32 with the step's gi
//     prefetched before the wait, h kept in registers across steps, writes out[b,t,dir*H+j], the saved
//     gates (layout [t][dir][gate][j][b]: coalesced along b) and the next operand image.
// Gate math follows PyTorch: r,z = sigmoid, n = tanh(gi_n + r*(W_hn h + b_hn)),
// h' = (1-z)*n + z*h, h0 = 0.
#include "s2ag.h"
#include "gemm_umma.cuh"

namespace s2ag {
extern int g_last_gru_kernel;
namespace grup {

using namespace s2ag::umma;

constexpr int PBM = 128, HS = 16, NC = 48, PTHREADS = 256, HDR = 128;

struct Params {
  const float* gi;          // [B*T][2][3H]  (x W_ih^T + b_ih)
  const float* whh;         // direction 0; direction 1 at + whh_dstride
  long whh_dstride;
  const float* bhh;
  long bhh_dstride;
  float* out;               // [B][T][2H]
  float* gates;             // [T][2][4][H][B] or nullptr
  unsigned char* xchg;      // operand images: [group][parity][hi|lo][Kpad/8][128][16 B]
  unsigned int* cnt;        // one arrival counter per group, 32 words apart
  int B, T, H, Kpad, S, nbt, bt0, x3;  // nbt: batch tiles of the whole batch; bt0: first tile of this launch
  int dbg;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
// same without the wait: the destination registers are valid only after tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
// 32-byte (one full sector) global accesses: 8 consecutive floats, address 32-byte aligned
__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(p)
               : "memory");
}
__device__ __forceinline__ void ld_global_cg_v8(const float* p, float (&v)[8]) {  // L1 bypassed (data written by other SMs)
  asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void pack8(const float (&v)[8], uint4& hi, uint4& lo) {
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

// Forward kernel: 9 warps.  Warps 0-7 ("workers", thread = clip row x 8 hidden units) fetch the operand image
// slice by slice as the producing CTAs publish it (per-slice release/acquire flags, cp.async.cg + mbarrier
// arrive-on-completion) and run the gate-math epilogue; warp 8 issues the tcgen05.mma k-step of a slice as soon as
// that slice has landed, so only the LAST slice to arrive is on the critical path of a time step.
// 8 worker warps + 4 MMA-issuer warps
constexpr int FWD_THREADS = 384, FWD_HDR = 512, MAX_SLICES = 24, NISSUE = 4;
// The k-steps of a time step are spread round-robin over NACC independent TMEM accumulators (two per issuer thread)
// that the epilogue sums, so that the MMAs of a slice never wait for the accumulator of another one.
constexpr int NACC = 8;  // = 2 per MMA-issuer thread

// bring-up aid (s2ag_debug_flags bit 1): clock64 timeline of CTA (slice 0, tile 0, direction 0), 16 marks per step
__device__ long long g_gru_timeline[64 * 16];
__device__ long long g_gru_slices[64 * 3 * MAX_SLICES];
__device__ long long g_gru_ctas[MAX_SLICES * 8];  // globaltimer marks of every slice CTA of group 0 at step 10
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }  // per step: [flag seen | copy issued | landed][slice]
#define GRU_MARK(slot) do { if (dbg) g_gru_timeline[(s & 63) * 16 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(FWD_THREADS, 1) gru_persist_fwd_kernel(Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, bt = p.bt0 + blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, B = p.B, Kpad = p.Kpad, S = p.S;
  const int nchunk = Kpad >> 3;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t mma_bar = sbase;            // all MMAs of a step have completed
  const uint32_t tfree_bar = sbase + 8;      // the 8 worker warps have read the accumulator of the previous step
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 16);
  const uint32_t ready0 = sbase + 64;        // ready[kk]: slice kk of the operand image has landed (TMA complete_tx)
  unsigned char* w_hi = smem + FWD_HDR;                // [nchunk][48][16]
  unsigned char* w_lo = w_hi + nchunk * NC * 16;
  unsigned char* a_hi = w_lo + nchunk * NC * 16;       // [nchunk][128][16]
  const int a_half = nchunk * PBM * 16;
  const int group = dir * p.nbt + bt;
  unsigned* flags = p.cnt + (size_t)group * MAX_SLICES * 32;   // one flag per slice, 128 bytes apart
  unsigned char* img0 = p.xchg + (size_t)group * 2 * (2 * (size_t)a_half);

  if (tid == 0) {
    mbar_init(mma_bar, NISSUE);
    mbar_init(tfree_bar, 8);
    for (int i = 0; i < S; ++i) mbar_init(ready0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 16, 512);

  // ---- stationary W_hh slice: rows n = g*16 + jj  <->  W_hh[g*H + j0 + jj][k]
  const float* whh = p.whh + dir * p.whh_dstride;
  const int j0 = slice * HS;
  for (int idx = tid; idx < nchunk * NC; idx += FWD_THREADS) {
    const int n = idx % NC, kc = idx / NC;
    const int g = n / HS, j = j0 + n % HS;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kc * 8 + i;
      v[i] = (j < H && k < H) ? __ldg(whh + ((long)g * H + j) * H + k) : 0.f;
    }
    uint4 hi, lo;
    pack8(v, hi, lo);
    *reinterpret_cast<uint4*>(w_hi + (kc * NC + n) * 16) = hi;
    *reinterpret_cast<uint4*>(w_lo + (kc * NC + n) * 16) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool dbg_cta = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // provably warp-uniform: issuer descriptors in uniform registers
  if (warp_u >= 8) {
    // ================================ MMA issuers ================================
    // One elected lane of each of the NISSUE issuer warps takes the k-steps kk = iss, iss + NISSUE, ... (each waits
    // on its own slices' mbarriers: four wait/issue streams keep the tail after the last slice short) and
    // accumulates into its own two TMEM accumulators; each issuer commits to mma_bar (NISSUE arrivals).
    if (elect_one()) {
      const int iss = warp_u - 8;
      const bool dbg = dbg_cta && iss == 0;
      const uint32_t idesc = make_idesc(NC);
      const uint32_t a_lbo = PBM * 16, w_lbo = NC * 16;
      const uint64_t dah0 = make_desc(smem_u32(a_hi), a_lbo, 128), dal0 = make_desc(smem_u32(a_hi) + a_half, a_lbo, 128);
      const uint64_t dwh0 = make_desc(smem_u32(w_hi), w_lbo, 128), dwl0 = make_desc(smem_u32(w_lo), w_lbo, 128);
      const uint32_t acc0 = tmem_base + (uint32_t)(2 * iss * NC), acc1 = acc0 + NC;
      for (int s = 1; s < T; ++s) {
        const uint32_t par = (uint32_t)((s - 1) & 1);
        mbar_wait(tfree_bar, par);  // accumulators of step s-1 have been read by every worker warp
        tc_fence_after();
        GRU_MARK(8);
        uint32_t c = 0;  // MMAs issued by this thread in this step
        for (int kk = iss; kk < S; kk += NISSUE) {
          mbar_wait(ready0 + 8 * kk, par);
          if (dbg) g_gru_slices[((s & 63) * 3 + 2) * MAX_SLICES + kk] = clock64();
          tc_fence_after();
          // descriptors of k-step kk: the start-address field (16-byte units) advances by two k-chunks
          const uint64_t dah = dah0 + (uint64_t)(kk * ((2 * a_lbo) >> 4)), dal = dal0 + (uint64_t)(kk * ((2 * a_lbo) >> 4));
          const uint64_t dwh = dwh0 + (uint64_t)(kk * ((2 * w_lbo) >> 4)), dwl = dwl0 + (uint64_t)(kk * ((2 * w_lbo) >> 4));
          if (p.x3) {
            mma_bf16((c & 1) ? acc1 : acc0, dal, dwh, idesc, c >= 2 ? 1u : 0u); ++c;
            mma_bf16((c & 1) ? acc1 : acc0, dah, dwl, idesc, c >= 2 ? 1u : 0u); ++c;
          }
          mma_bf16((c & 1) ? acc1 : acc0, dah, dwh, idesc, c >= 2 ? 1u : 0u); ++c;
        }
        mma_commit(mma_bar);
        GRU_MARK(11);
      }
    }
  } else {
    // ================================ workers ================================
    const int row = (warp & 3) * 32 + lane;
    const int b = bt * PBM + row;
    const int u0 = (warp >> 2) * 8;          // unit offset inside the slice
    const int jb = j0 + u0;                  // first hidden unit of this thread
    const bool b_ok = b < B;
    // 16-byte accesses to gi / out rows: all 8 units valid and every row segment 16-byte aligned
    const bool vec_ok = (H & 3) == 0 &&
                        ((reinterpret_cast<uintptr_t>(p.gi) | reinterpret_cast<uintptr_t>(p.out)) & 15) == 0;
    // the last slice may own a partial group of units: validity per 4 units (H % 4 == 0 on this path)
    const bool q0_ok = jb + 4 <= H, q1_ok = jb + 8 <= H;
    const float* bhh = p.bhh + dir * p.bhh_dstride;
    float br[8], bz[8], bn[8], h_own[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = jb + i < H;
      br[i] = ok ? __ldg(bhh + jb + i) : 0.f;
      bz[i] = ok ? __ldg(bhh + H + jb + i) : 0.f;
      bn[i] = ok ? __ldg(bhh + 2 * H + jb + i) : 0.f;
      h_own[i] = 0.f;
    }
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // accumulator a = 2*issuer + j is written in a step iff that issuer issues more than j MMAs
    unsigned acc_mask = 0;
#pragma unroll
    for (int i = 0; i < NISSUE; ++i) {
      const int nk = i < S ? (S - i + NISSUE - 1) / NISSUE : 0;
      const int nm = nk * (p.x3 ? 3 : 1);
      if (nm > 0) acc_mask |= 1u << (2 * i);
      if (nm > 1) acc_mask |= 2u << (2 * i);
    }
    const bool dbg = dbg_cta && tid == 0;
    float ph[8], pr[8], pz[8], pn[8], pg[8];
    int pt = -1;
    auto store_prev = [&]() {
      if (pt < 0 || !b_ok) { pt = -1; return; }
      float* orow = p.out + ((long)b * T + pt) * 2 * H + (long)dir * H + jb;
      if (vec_ok) {
        if (q0_ok) reinterpret_cast<float4*>(orow)[0] = make_float4(ph[0], ph[1], ph[2], ph[3]);
        if (q1_ok) reinterpret_cast<float4*>(orow)[1] = make_float4(ph[4], ph[5], ph[6], ph[7]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (jb + i < H) orow[i] = ph[i];
      }
      if (p.gates) {
        float* gs = p.gates + ((((long)pt * 2 + dir) * 4) * H + jb) * B + b;
        const long gstride = (long)H * B;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (jb + i < H) {
            gs[(long)i * B] = pr[i]; gs[gstride + (long)i * B] = pz[i]; gs[2 * gstride + (long)i * B] = pn[i];
            gs[3 * gstride + (long)i * B] = pg[i];
          }
        }
      }
      pt = -1;
    };

    for (int s = 0; s < T; ++s) {
      const int t = dir == 0 ? s : T - 1 - s;
      GRU_MARK(0);
      // gi of this step: independent of the recurrence, issue the loads first
      float gr[8], gz[8], gn[8];
      {
        const float* g = p.gi + ((long)(b_ok ? b : 0) * T + t) * 6 * H + (long)dir * 3 * H + jb;
        if (vec_ok) {
          if (b_ok) {
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 r0 = q0_ok ? __ldg(reinterpret_cast<const float4*>(g)) : zero4;
            const float4 r1 = q1_ok ? __ldg(reinterpret_cast<const float4*>(g) + 1) : zero4;
            const float4 z0 = q0_ok ? __ldg(reinterpret_cast<const float4*>(g + H)) : zero4;
            const float4 z1 = q1_ok ? __ldg(reinterpret_cast<const float4*>(g + H) + 1) : zero4;
            const float4 n0 = q0_ok ? __ldg(reinterpret_cast<const float4*>(g + 2 * H)) : zero4;
            const float4 n1 = q1_ok ? __ldg(reinterpret_cast<const float4*>(g + 2 * H) + 1) : zero4;
            gr[0] = r0.x; gr[1] = r0.y; gr[2] = r0.z; gr[3] = r0.w; gr[4] = r1.x; gr[5] = r1.y; gr[6] = r1.z; gr[7] = r1.w;
            gz[0] = z0.x; gz[1] = z0.y; gz[2] = z0.z; gz[3] = z0.w; gz[4] = z1.x; gz[5] = z1.y; gz[6] = z1.z; gz[7] = z1.w;
            gn[0] = n0.x; gn[1] = n0.y; gn[2] = n0.z; gn[3] = n0.w; gn[4] = n1.x; gn[5] = n1.y; gn[6] = n1.z; gn[7] = n1.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) gr[i] = gz[i] = gn[i] = 0.f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = b_ok && jb + i < H;
            gr[i] = ok ? __ldg(g + i) : 0.f;
            gz[i] = ok ? __ldg(g + H + i) : 0.f;
            gn[i] = ok ? __ldg(g + 2 * H + i) : 0.f;
          }
        }
      }
      GRU_MARK(14);
      float ar[8], az[8], an[8];
      if (s > 0) {
        // fetch the operand image slice by slice as soon as each producer has published h_{s-1}: lanes 0..2 of every
        // worker warp own one slice each (rotated by the CTA's slice index), poll its flag and launch TMA bulk copies
        // (hi and lo k-chunk pairs, 4096 contiguous bytes each) that complete on the slice's mbarrier.
        {
          const int my_i = warp + 8 * lane;
          if (lane < 3 && my_i < S) {
            const int sl = (my_i + slice) % S;
            unsigned spins = 0;
            while (ld_acquire_u32(flags + sl * 32) < (unsigned)s) {
              if (++spins > (1u << 26)) __trap();
            }
            GRU_MARK(12);
            GRU_MARK(1);
            if (dbg_cta) g_gru_slices[((s & 63) * 3 + 0) * MAX_SLICES + sl] = clock64();
            asm volatile("fence.proxy.async.global;" ::: "memory");  // generic-proxy writes (other SMs) -> async-proxy read
            GRU_MARK(13);
            const unsigned char* src = img0 + (size_t)((s - 1) & 1) * 2 * a_half + (size_t)sl * 2 * PBM * 16;
            const uint32_t dst = smem_u32(a_hi) + (uint32_t)(sl * 2 * PBM * 16);
            const uint32_t bar = ready0 + 8 * sl;
            const uint32_t nbytes = 2 * PBM * 16;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(p.x3 ? 2 * nbytes : nbytes)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src), "r"(nbytes), "r"(bar)
                         : "memory");
            if (p.x3)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                               dst + (uint32_t)a_half),
                           "l"(src + a_half), "r"(nbytes), "r"(bar)
                           : "memory");
            if (dbg_cta) g_gru_slices[((s & 63) * 3 + 1) * MAX_SLICES + sl] = clock64();
          }
          __syncwarp();
        }
        if (p.dbg && group == 0 && s == 10 && tid == 0) g_gru_ctas[slice * 8 + 1] = gtimer();
        // layer output and saved gates of the PREVIOUS step: issued while the image is in flight
        store_prev();
        if (p.dbg && group == 0 && s == 10 && tid == 0) g_gru_ctas[slice * 8 + 2] = gtimer();
        GRU_MARK(2);
        mbar_wait(mma_bar, (uint32_t)((s - 1) & 1));
        GRU_MARK(3);
        if (p.dbg && group == 0 && s == 10 && tid == 0) g_gru_ctas[slice * 8 + 3] = gtimer();
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < 8; ++i) ar[i] = az[i] = an[i] = 0.f;
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          if (acc_mask & (1u << a)) {
            float v0[8], v1[8], v2[8];
            tmem_ld8_nowait(t_lane + (uint32_t)(a * NC + 0 * HS + u0), v0);
            tmem_ld8_nowait(t_lane + (uint32_t)(a * NC + 1 * HS + u0), v1);
            tmem_ld8_nowait(t_lane + (uint32_t)(a * NC + 2 * HS + u0), v2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) { ar[i] += v0[i]; az[i] += v1[i]; an[i] += v2[i]; }
          }
        }
        GRU_MARK(4);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) ar[i] = az[i] = an[i] = 0.f;
      }
      // the accumulator may be overwritten by the next step's MMAs once every worker warp got here
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && s + 1 < T) mbar_arrive(tfree_bar);
      // gate math (fast exp / reciprocal: |error| ~1e-7, far inside the parity budget)
      float rr[8], zz[8], nn[8], gh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = b_ok && jb + i < H;
        gh[i] = an[i] + bn[i];
        rr[i] = __fdividef(1.f, 1.f + __expf(-(gr[i] + ar[i] + br[i])));
        zz[i] = __fdividef(1.f, 1.f + __expf(-(gz[i] + az[i] + bz[i])));
        nn[i] = 1.f - __fdividef(2.f, __expf(2.f * (gn[i] + rr[i] * gh[i])) + 1.f);
        const float h = (1.f - zz[i]) * nn[i] + zz[i] * h_own[i];
        h_own[i] = ok ? h : 0.f;
      }
      if (s + 1 < T) {
        // publish the operand image of h_s (parity s&1): chunk (jb/8), row `row`; raise this slice's flag as soon as
        // the image stores of the 8 worker warps are visible (the out / gates stores below are not waited for)
        uint4 hi, lo;
        pack8(h_own, hi, lo);
        unsigned char* img = img0 + (size_t)(s & 1) * 2 * a_half;
        const int off = ((jb >> 3) * PBM + row) * 16;
        *reinterpret_cast<uint4*>(img + off) = hi;
        if (p.x3) *reinterpret_cast<uint4*>(img + a_half + off) = lo;
        GRU_MARK(5);
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 worker warps
        GRU_MARK(6);
        if (p.dbg && group == 0 && s == 10 && tid == 0) g_gru_ctas[slice * 8 + 4] = gtimer();
        // release: cumulative over bar.sync.  (The gpu-scope fence costs ~1.3k cycles and stalls the memory
        // instructions of every warp of the SM meanwhile -- measured; moving it to a helper warp gains nothing.)
        if (tid == 0) st_release_u32(flags + slice * 32, (unsigned)(s + 1));
        GRU_MARK(7);
        if (p.dbg && group == 0 && s == 10 && tid == 0) g_gru_ctas[slice * 8 + 5] = gtimer();
        if (p.dbg && group == 0 && s == 9 && tid == 0) g_gru_ctas[slice * 8 + 0] = gtimer();
      }
      // layer output and saved gates: kept in registers, stored at the next step while its image is in flight
#pragma unroll
      for (int i = 0; i < 8; ++i) { ph[i] = h_own[i]; pr[i] = rr[i]; pz[i] = zz[i]; pn[i] = nn[i]; pg[i] = gh[i]; }
      pt = t;
      if (s == 0 || s + 1 == T) store_prev();  // step 0 has no fetch phase; the last step has no successor
      GRU_MARK(15);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------
// BPTT.  Same ownership (CTA = 16 hidden units x 128 clips x direction, resident for all T steps).  Per step
// (forward order reversed) the CTA computes the gate gradients of its units locally,
//   dh = dout + carry;  dn = dh(1-z)(1-n^2);  dz = dh(h_prev - n)z(1-z);  dr = dn*ghn*r(1-r),
// writes dgi = (dr,dz,dn) and dgh = (dr,dz,dn*r) for the time-batched weight-gradient GEMMs, and contributes
//   partial[b, k] = sum_{c in its 48 gate rows} dgh[b,c] * W_hh[c, k]          (all k: M=128, N=Kpad, K=48)
// on tcgen05 (A = its own dgh, split hi/lo from registers; B = its 48 rows of W_hh, stationary, stored as
// [k-chunk of c][n = k][16 B]).  The partials ([slice][k][b], coalesced along b) are exchanged through L2; after
// the per-(tile, direction) arrival counter each CTA sums the 16 columns it owns over all slices:
//   carry[b, j] = dh*z + sum_slices partial_s[b, j].
struct BwdParams {
  const float* dout; long lddout; int dir_stride;
  const float* out;         // [B][T][2H] forward output (h_prev)
  const float* gates;       // [T][2][4][H][B]
  const float* whh; long whh_dstride;
  float* dgi;               // [B*T][2][3H]
  float* dgh;               // [B*T][2][3H]
  float* part;              // [group][parity][S][Kpad][128]
  unsigned int* cnt;
  int B, T, H, Kpad, S, nbt, bt0, x3;
  int dbg;
};

__global__ void __launch_bounds__(PTHREADS, 1) gru_persist_bwd_kernel(BwdParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, bt = p.bt0 + blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, B = p.B, Kpad = p.Kpad, S = p.S;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t mma_bar = sbase;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 8);
  // B operand: [6 k-chunks][Kpad rows n][16 B] hi, then lo ; A operand: [6 k-chunks][128][16 B] hi, then lo
  unsigned char* w_hi = smem + HDR;
  const int w_half = 6 * Kpad * 16;
  unsigned char* w_lo = w_hi + w_half;
  unsigned char* a_hi = w_lo + w_half;
  const int a_half = 6 * PBM * 16;
  unsigned char* a_lo = a_hi + a_half;
  const int group = dir * p.nbt + bt;
  const unsigned* cnt = p.cnt + group * 32;
  const size_t part_slice = (size_t)Kpad * PBM;            // floats of one slice's partial
  float* part0 = p.part + (size_t)group * 2 * S * part_slice;
  const uint32_t ncols = Kpad <= 32 ? 32u : Kpad <= 64 ? 64u : Kpad <= 128 ? 128u : Kpad <= 256 ? 256u : 512u;

  if (tid == 0) {
    mbar_init(mma_bar, Kpad <= 256 ? 1 : 2);  // one arrival per MMA-issuer thread (one per accumulator half)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 8, ncols);

  // ---- stationary W_hh rows of this slice, transposed for the B operand: element (n = k_out, kk = c_local)
  const float* whh = p.whh + dir * p.whh_dstride;
  const int j0 = slice * HS;
  for (int idx = tid; idx < 6 * Kpad; idx += PTHREADS) {
    const int n = idx % Kpad, kc = idx / Kpad;       // kc: chunk of 8 local gate rows
    const int g = kc >> 1, jj0 = (kc & 1) * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = j0 + jj0 + i;
      v[i] = (j < H && n < H) ? __ldg(whh + ((long)g * H + j) * H + n) : 0.f;
    }
    uint4 hi, lo;
    pack8(v, hi, lo);
    *reinterpret_cast<uint4*>(w_hi + (kc * Kpad + n) * 16) = hi;
    *reinterpret_cast<uint4*>(w_lo + (kc * Kpad + n) * 16) = lo;
  }

  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // provably warp-uniform: issuer descriptors in uniform registers
  const int row = (warp & 3) * 32 + lane;
  const int b = bt * PBM + row;
  const int wg = warp >> 2;
  const int u0 = wg * 8, jb = j0 + u0;
  const bool b_ok = b < B;
  const bool vec_ok = (H & 3) == 0 && (p.lddout & 3) == 0 && ((p.dir_stride & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.dout) | reinterpret_cast<uintptr_t>(p.out) |
                        reinterpret_cast<uintptr_t>(p.dgi) | reinterpret_cast<uintptr_t>(p.dgh)) & 15) == 0;
  // 32-byte stores of the dgi / dgh rows: every gate segment (3H*4, H*4, jb*4 bytes) must keep 32-byte alignment
  const bool vec32_ok = vec_ok && (H & 7) == 0 &&
                        ((reinterpret_cast<uintptr_t>(p.dgi) | reinterpret_cast<uintptr_t>(p.dgh)) & 31) == 0;
  // the last slice may own a partial group of units: validity per 4 units (H % 4 == 0 on the vector paths)
  const bool q0_ok = jb + 4 <= H, q1_ok = jb + 8 <= H;
  float carry[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) carry[i] = 0.f;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  // N split: one MMA covers at most 256 accumulator columns
  const int n1 = Kpad <= 256 ? Kpad : 160, n2 = Kpad - n1;

  // per-step operands (dout, h_prev, saved gates): loaded one step ahead, they do not depend on the recurrence
  float ndh[8], nhp[8], nr[8], nz[8], nn_[8], nghn[8];
  auto load_step = [&](int step) {
    const int fs = T - 1 - step;
    const int t = dir == 0 ? fs : T - 1 - fs;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const long rowi = (long)(b_ok ? b : 0) * T + t;
    const float* dp = p.dout + rowi * p.lddout + (long)dir * p.dir_stride + jb;
    const float* hq = p.out + ((long)(b_ok ? b : 0) * T + (fs > 0 ? tprev : t)) * 2 * H + (long)dir * H + jb;
    if (vec_ok && b_ok) {
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 d0 = q0_ok ? __ldg(reinterpret_cast<const float4*>(dp)) : zero4;
      const float4 d1 = q1_ok ? __ldg(reinterpret_cast<const float4*>(dp) + 1) : zero4;
      ndh[0] = d0.x; ndh[1] = d0.y; ndh[2] = d0.z; ndh[3] = d0.w; ndh[4] = d1.x; ndh[5] = d1.y; ndh[6] = d1.z; ndh[7] = d1.w;
      if (fs > 0) {
        const float4 h0 = q0_ok ? __ldg(reinterpret_cast<const float4*>(hq)) : zero4;
        const float4 h1 = q1_ok ? __ldg(reinterpret_cast<const float4*>(hq) + 1) : zero4;
        nhp[0] = h0.x; nhp[1] = h0.y; nhp[2] = h0.z; nhp[3] = h0.w; nhp[4] = h1.x; nhp[5] = h1.y; nhp[6] = h1.z; nhp[7] = h1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) nhp[i] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = b_ok && jb + i < H;
        ndh[i] = ok ? __ldg(dp + i) : 0.f;
        nhp[i] = (ok && fs > 0) ? __ldg(hq + i) : 0.f;
      }
    }
    const float* gs = p.gates + ((((long)t * 2 + dir) * 4) * H + jb) * B + (b_ok ? b : 0);
    const long gstride = (long)H * B;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = b_ok && jb + i < H;
      nr[i] = ok ? __ldg(gs + (long)i * B) : 0.f;
      nz[i] = ok ? __ldg(gs + gstride + (long)i * B) : 0.f;
      nn_[i] = ok ? __ldg(gs + 2 * gstride + (long)i * B) : 0.f;
      nghn[i] = ok ? __ldg(gs + 3 * gstride + (long)i * B) : 0.f;
    }
  };
  load_step(0);
  const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
#define BWD_MARK(slot) do { if (dbg) g_gru_timeline[(step & 63) * 16 + (slot)] = clock64(); } while (0)

  for (int step = 0; step < T; ++step) {
    BWD_MARK(0);
    const int fs = T - 1 - step;                         // forward step index being differentiated
    const int t = dir == 0 ? fs : T - 1 - fs;
    const long rowi = (long)(b_ok ? b : 0) * T + t;
    float dh[8], hp[8], r[8], z[8], n[8], ghn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dh[i] = ndh[i]; hp[i] = nhp[i]; r[i] = nr[i]; z[i] = nz[i]; n[i] = nn_[i]; ghn[i] = nghn[i]; }
    if (step + 1 < T) load_step(step + 1);
    float dr[8], dz[8], dn[8], dnr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dh[i] += carry[i];
      dn[i] = dh[i] * (1.f - z[i]) * (1.f - n[i] * n[i]);
      dz[i] = dh[i] * (hp[i] - n[i]) * z[i] * (1.f - z[i]);
      dr[i] = dn[i] * ghn[i] * r[i] * (1.f - r[i]);
      dnr[i] = dn[i] * r[i];
      carry[i] = dh[i] * z[i];
    }
    auto store_grads = [&]() {  // dgi / dgh rows for the time-batched weight-gradient GEMMs
      if (!b_ok) return;
      float* a = p.dgi + (rowi * 2 + dir) * 3 * H + jb;
      float* c = p.dgh + (rowi * 2 + dir) * 3 * H + jb;
      if (vec32_ok) {  // H % 8 == 0: a group of 8 units is valid as a whole or not at all
        if (q1_ok) {
          st_global_v8(a, dr); st_global_v8(c, dr);
          st_global_v8(a + H, dz); st_global_v8(c + H, dz);
          st_global_v8(a + 2 * H, dn); st_global_v8(c + 2 * H, dnr);
        }
      } else if (vec_ok) {
        auto st2 = [&](float* dst, const float (&v)[8]) {
          if (q0_ok) reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
          if (q1_ok) reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
        };
        st2(a, dr); st2(c, dr);
        st2(a + H, dz); st2(c + H, dz);
        st2(a + 2 * H, dn); st2(c + 2 * H, dnr);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (jb + i < H) {
            a[i] = dr[i]; a[H + i] = dz[i]; a[2 * H + i] = dn[i];
            c[i] = dr[i]; c[H + i] = dz[i]; c[2 * H + i] = dnr[i];
          }
        }
      }
    };
    if (fs == 0) store_grads();
    if (fs > 0) {
      // A operand = this CTA's dgh tile [128 x 48]: local gate row c = g*16 + u0 + i  ->  chunk g*2 + wg
      {
        uint4 hi, lo;
        pack8(dr, hi, lo);
        *reinterpret_cast<uint4*>(a_hi + ((0 * 2 + wg) * PBM + row) * 16) = hi;
        *reinterpret_cast<uint4*>(a_lo + ((0 * 2 + wg) * PBM + row) * 16) = lo;
        pack8(dz, hi, lo);
        *reinterpret_cast<uint4*>(a_hi + ((1 * 2 + wg) * PBM + row) * 16) = hi;
        *reinterpret_cast<uint4*>(a_lo + ((1 * 2 + wg) * PBM + row) * 16) = lo;
        pack8(dnr, hi, lo);
        *reinterpret_cast<uint4*>(a_hi + ((2 * 2 + wg) * PBM + row) * 16) = hi;
        *reinterpret_cast<uint4*>(a_lo + ((2 * 2 + wg) * PBM + row) * 16) = lo;
      }
      fence_async_smem();
      BWD_MARK(1);
      __syncthreads();
      BWD_MARK(2);
      if ((warp_u == 0 || (warp_u == 1 && n2 > 0)) && elect_one()) {
        // one issuer thread per accumulator half: 9 dependent tcgen05.mma each instead of 18 in one stream
        tc_fence_after();
        const int half = warp_u;
        const uint32_t sa = smem_u32(a_hi), sw = smem_u32(w_hi);
        const uint32_t a_lbo = PBM * 16, w_lbo = (uint32_t)Kpad * 16;
        const int nn = half == 0 ? n1 : n2;
        const uint32_t idesc = make_idesc(nn);
        const uint32_t d_tmem = tmem_base + (half == 0 ? 0u : (uint32_t)n1);
        const uint32_t w_off = half == 0 ? 0u : (uint32_t)n1 * 16u;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const uint32_t ah = sa + kk * 2 * a_lbo, al = ah + a_half;
          const uint32_t wh = sw + kk * 2 * w_lbo + w_off, wl = wh + w_half;
          const uint64_t dah = make_desc(ah, a_lbo, 128), dwh = make_desc(wh, w_lbo, 128);
          uint32_t acc = kk > 0 ? 1u : 0u;
          if (p.x3) {
            mma_bf16(d_tmem, make_desc(al, a_lbo, 128), dwh, idesc, acc);
            mma_bf16(d_tmem, dah, make_desc(wl, w_lbo, 128), idesc, 1u);
            acc = 1u;
          }
          mma_bf16(d_tmem, dah, dwh, idesc, acc);
        }
        mma_commit(mma_bar);
      }
      mbar_wait(mma_bar, (uint32_t)(step & 1));
      BWD_MARK(3);
      tc_fence_after();
      // accumulator -> partial[parity][slice][k][b]; warpgroup wg takes the k-columns [wg*Kpad/2, +Kpad/2)
      float* mypart = part0 + ((size_t)(step & 1) * S + slice) * part_slice;
      const int kh = Kpad >> 1;
      for (int c0 = wg * kh; c0 < (wg + 1) * kh; c0 += 8) {
        float v[8];
        tmem_ld8(t_lane + (uint32_t)c0, v);
        st_global_v8(mypart + (size_t)row * Kpad + c0, v);  // partial[slice][b][k]: one 32-byte sector per store
      }
      tc_fence_before();
      BWD_MARK(4);
      __syncthreads();
      BWD_MARK(5);
      if (tid == 0) {
        // arrive (release: cumulative over the __syncthreads above) and wait for the other slices of the group
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
      }
      BWD_MARK(6);
      store_grads();  // off the critical path: overlaps the wait for the other CTAs
      BWD_MARK(7);
      if (tid == 0) {
        const unsigned target = (unsigned)S * (unsigned)(step + 1);
        unsigned spins = 0;
        while (*reinterpret_cast<const volatile unsigned*>(cnt) < target) {
          if (++spins > (1u << 26)) __trap();
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      BWD_MARK(8);
      __syncthreads();
      BWD_MARK(9);
      // carry[b, j] += sum over slices of their partial columns j (this thread: 8 columns, its clip)
      const float* pp = part0 + (size_t)(step & 1) * S * part_slice + (size_t)row * Kpad + jb;
      for (int sl = 0; sl < S; ++sl) {
        float v[8];
        ld_global_cg_v8(pp + (size_t)sl * part_slice, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) carry[i] += v[i];
      }
      BWD_MARK(10);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

static inline int kpad_of(int H) { return ((H + HS - 1) / HS) * HS; }
static inline size_t img_bytes(int H) { return 2 * (size_t)(kpad_of(H) / 8) * PBM * 16; }  // hi + lo
static inline size_t persist_smem_bytes(int H) { return FWD_HDR + (size_t)kpad_of(H) / 8 * (2 * NC + 2 * PBM) * 16; }

}  // namespace grup

int gru_debug_read_timeline(long long* host, int n) {
  // n <= 1024: step marks; n > 1024: step marks followed by the per-slice table
  int n1 = n < 64 * 16 ? n : 64 * 16;
  if (cudaMemcpyFromSymbol(host, grup::g_gru_timeline, sizeof(long long) * n1) != cudaSuccess) return -2;
  int n2 = n - n1;
  if (n2 > 64 * 3 * grup::MAX_SLICES) n2 = 64 * 3 * grup::MAX_SLICES;
  if (n2 > 0 && cudaMemcpyFromSymbol(host + n1, grup::g_gru_slices, sizeof(long long) * n2) != cudaSuccess) return -2;
  int n3 = n - n1 - 64 * 3 * grup::MAX_SLICES;
  if (n3 > grup::MAX_SLICES * 8) n3 = grup::MAX_SLICES * 8;
  if (n3 > 0 && cudaMemcpyFromSymbol(host + n1 + 64 * 3 * grup::MAX_SLICES, grup::g_gru_ctas, sizeof(long long) * n3) != cudaSuccess) return -2;
  return 0;
}

// A launch keeps every CTA resident: at most 148 CTAs -> batch chunks of `rows_per_launch` clips.
static size_t bwd_smem_bytes(int H);
static bool gru_persist_resident(int H);
bool gru_persist_supported(int H) {
  return H >= 16 && grup::persist_smem_bytes(H) <= 227 * 1024 && grup::kpad_of(H) / grup::HS <= grup::MAX_SLICES &&
         gru_persist_resident(H);
}
static int gru_persist_tiles_per_launch(int H) {
  const int S = grup::kpad_of(H) / grup::HS;
  int tiles = s2ag_sm_count() / (2 * S);   // every CTA of a launch must be resident (inter-CTA flags)
  return tiles < 1 ? 0 : tiles;
}
// Residency is verified, not assumed: the CTAs of a launch poll each other's flags, so all S x 2 x tiles CTAs of the
// largest launch must fit on the device at once (occupancy x SMs); otherwise gru_persist_supported() is false and the
// per-step kernels of gru.cu run.  Cached per hidden size.
static bool gru_persist_resident(int H) {
  static int cache[512];   // 0 unknown, 1 yes, 2 no
  if (H < 0 || H >= 512) return false;
  if (cache[H]) return cache[H] == 1;
  using namespace grup;
  const int S = kpad_of(H) / HS, tiles = gru_persist_tiles_per_launch(H);
  bool ok = tiles > 0;
  if (ok) {
    cudaFuncSetAttribute(&gru_persist_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(&gru_persist_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    int f = 0, b = 0;
    ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&f, &gru_persist_fwd_kernel, FWD_THREADS, persist_smem_bytes(H)) == cudaSuccess &&
         cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, &gru_persist_bwd_kernel, PTHREADS, bwd_smem_bytes(H)) == cudaSuccess &&
         (long)f * s2ag_sm_count() >= (long)S * 2 * tiles && (long)b * s2ag_sm_count() >= (long)S * 2 * tiles;
    if (!ok) cudaGetLastError();
  }
  cache[H] = ok ? 1 : 2;
  return ok;
}
// bytes of exchange workspace (operand images + counters) for a batch of B clips
size_t gru_persist_ws_bytes(int B, int H) {
  if (!gru_persist_supported(H) || gru_persist_tiles_per_launch(H) == 0) return 0;
  const int nbt = (B + grup::PBM - 1) / grup::PBM;
  return (size_t)2 * nbt * 2 * grup::img_bytes(H) + (size_t)2 * nbt * grup::MAX_SLICES * 32 * sizeof(unsigned) + 512;
}

int gru_persist_fwd(const float* gi, const float* whh_f, long whh_dstride, const float* bhh_f, long bhh_dstride,
                    float* out, float* gates, void* ws, int B, int T, int H, int x3, void* stream) {
  using namespace grup;
  const int S = kpad_of(H) / HS;
  const int tiles_max = gru_persist_tiles_per_launch(H);
  if (tiles_max == 0 || !gru_persist_supported(H)) return S2AG_ERR_UNSUPPORTED;
  auto kfn = &gru_persist_fwd_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const int nbt_all = (B + PBM - 1) / PBM;
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  base += (256 - (reinterpret_cast<uintptr_t>(base) & 255)) & 255;
  unsigned int* cnt = reinterpret_cast<unsigned int*>(base);
  const size_t flag_bytes = (size_t)2 * nbt_all * MAX_SLICES * 32 * sizeof(unsigned);  // [group][slice], 128 B apart
  unsigned char* xchg = base + ((flag_bytes + 255) & ~(size_t)255);
  if (cudaMemsetAsync(cnt, 0, flag_bytes, (cudaStream_t)stream) != cudaSuccess) return S2AG_ERR_LAUNCH;
  for (int t0 = 0; t0 < nbt_all; t0 += tiles_max) {  // every launch keeps all of its CTAs resident
    const int nbt = nbt_all - t0 < tiles_max ? nbt_all - t0 : tiles_max;
    Params p;
    p.gi = gi; p.whh = whh_f; p.whh_dstride = whh_dstride; p.bhh = bhh_f; p.bhh_dstride = bhh_dstride;
    p.out = out; p.gates = gates; p.xchg = xchg; p.cnt = cnt;
    p.B = B; p.T = T; p.H = H; p.Kpad = kpad_of(H); p.S = S; p.nbt = nbt_all; p.bt0 = t0; p.x3 = x3;
    p.dbg = (umma::g_dbg_flags & 2) ? 1 : 0;
    g_last_gru_kernel = 0;
    S2AG_LAUNCH(kfn, dim3(S, nbt, 2), FWD_THREADS, persist_smem_bytes(H), stream, p);
  }
  return S2AG_OK;
}


static size_t bwd_smem_bytes(int H) { return grup::HDR + (size_t)2 * 6 * 16 * (grup::kpad_of(H) + grup::PBM); }
size_t gru_persist_bwd_ws_bytes(int B, int H) {
  if (!gru_persist_supported(H) || gru_persist_tiles_per_launch(H) == 0) return 0;
  if (grup::kpad_of(H) > 512) return 0;
  const int nbt = (B + grup::PBM - 1) / grup::PBM;
  const int S = grup::kpad_of(H) / grup::HS;
  return (size_t)2 * nbt * 2 * S * grup::kpad_of(H) * grup::PBM * sizeof(float) + (size_t)2 * nbt * 32 * sizeof(unsigned) + 512;
}

int gru_persist_bwd(const float* dout, long lddout, int dir_stride, const float* out, const float* gates,
                    const float* whh_f, long whh_dstride, float* dgi, float* dgh, void* ws, int B, int T, int H, int x3,
                    void* stream) {
  using namespace grup;
  const int S = kpad_of(H) / HS;
  const int tiles_max = gru_persist_tiles_per_launch(H);
  if (tiles_max == 0 || gru_persist_bwd_ws_bytes(B, H) == 0) return S2AG_ERR_UNSUPPORTED;
  auto kfn = &gru_persist_bwd_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const int nbt_all = (B + PBM - 1) / PBM;
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  base += (256 - (reinterpret_cast<uintptr_t>(base) & 255)) & 255;
  unsigned int* cnt = reinterpret_cast<unsigned int*>(base);
  float* part = reinterpret_cast<float*>(base + (((size_t)2 * nbt_all * 32 * sizeof(unsigned) + 255) & ~(size_t)255));
  if (cudaMemsetAsync(cnt, 0, (size_t)2 * nbt_all * 32 * sizeof(unsigned), (cudaStream_t)stream) != cudaSuccess)
    return S2AG_ERR_LAUNCH;
  for (int t0 = 0; t0 < nbt_all; t0 += tiles_max) {
    const int nbt = nbt_all - t0 < tiles_max ? nbt_all - t0 : tiles_max;
    BwdParams p;
    p.dout = dout; p.lddout = lddout; p.dir_stride = dir_stride; p.out = out; p.gates = gates;
    p.whh = whh_f; p.whh_dstride = whh_dstride; p.dgi = dgi; p.dgh = dgh; p.part = part; p.cnt = cnt;
    p.B = B; p.T = T; p.H = H; p.Kpad = kpad_of(H); p.S = S; p.nbt = nbt_all; p.bt0 = t0; p.x3 = x3;
    p.dbg = (umma::g_dbg_flags & 8) ? 1 : 0;
    S2AG_LAUNCH(kfn, dim3(S, nbt, 2), PTHREADS, bwd_smem_bytes(H), stream, p);
  }
  return S2AG_OK;
}

}  // namespace s2ag
