// tcgen05 (UMMA) tiled contraction  C[m,n] = epi( sum_k A(m,k) * B(n,k) )  for sm_100a, with the same
// pluggable operand loaders / epilogues as the SIMT engine in gemm_simt.cuh, which it replaces
// for every dense shape of the hot path (GRU input projections, TCN conv-as-GEMM, Conv1d/2d
// implicit GEMM, linear heads, weight-gradient contractions).
//
// Numerics.  The reference computes in fp32 and the parity bar is 1e-3 relative on losses and
// poses (1e-4 in the module tests), which single-pass bf16 does not meet (4.5e-3).  The
// operands therefore stay fp32 in HBM and are split ON THE FLY while they are staged into shared
// memory:  x = hi + lo,  hi = bf16(x),  lo = bf16(x - hi)  (x - hi is exact in fp32), and the
// tensor core accumulates  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  in fp32 in TMEM ("bf16x3", ~2^-17
// relative operand error).  precision mode 1 ("bf16x1") drops the two correction terms.
//
// Structure of one CTA (544 threads, one 128 x BN output tile, BN <= 256 a multiple of 16):
//   - 16 staging warps: global -> registers (two ping-pong register sets: the loads of the next k-block
//     are in flight while the current one is converted) -> bf16 hi/lo -> shared memory in the canonical
//     no-swizzle K-major UMMA layout [k-chunk of 8][row][16 B], two stages; each warp fences
//     (fence.proxy.async) and arrives on the stage's "full" mbarrier;
//   - a 17th warp's elected thread waits for "full", issues the tcgen05.mma instructions of the stage
//     (M=128, N=BN, K=16 each; the issuer branch is warp-uniform -- __shfl_sync'ed warp index + elect.sync -- so
//     that the descriptors live in uniform registers and UTCHMMAs issue back to back; under a `lane == 0` branch
//     the compiler wraps every tcgen05.mma in an elect / R2UR.BROADCAST loop of ~160 cycles) and
//     tcgen05.commit's the stage's "empty" mbarrier: the
//     tensor core works on stage s while the warps stage s^1, with no block-wide barrier in the loop;
//   - accumulator: BN fp32 columns x 128 lanes of TMEM; epilogue: tcgen05.ld (32 lanes x 32
//     columns per warp), transposed through shared memory so that global stores (and the
//     epilogue's bias / residual loads) are coalesced along n.
// grid = (tilesN, tilesM, nbatch*splitk) exactly like the SIMT engine.
#pragma once
#include "gemm_simt.cuh"
#ifndef S2AG_EMU
#include <cuda_bf16.h>

namespace s2ag {
namespace umma {

constexpr int BM = 128, BK = 32, STAGE_THREADS = 512, THREADS = 544, BN_MAX = 256, EPI_WARPS = 16;  // 16 staging warps + 1 MMA-issuer warp
constexpr int A_STAGE_BYTES = BM * BK * 2 * 2;  // hi + lo, bf16
constexpr int HEADER_BYTES = 128;               // mbarriers + TMEM base slot

extern int g_precision;   // 0 = bf16x3 (fp32-grade, default), 1 = bf16x1
extern int g_dbg_flags;   // reserved for bring-up experiments

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// bounded spin: a protocol bug traps (launch error) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of a converged warp (the branch around it must be warp-uniform)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 inputs, fp32 accumulate, M=128, N and majors from idesc
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread `lane` receives row (lane_base + lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle.  Canonical layout (16-byte units):
// ((8 rows, n groups), 2 k-chunks) : ((1, SBO), LBO)  -- cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::K>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version 1 (sm_100); layout_type (bits 61-63) = 0: no swizzle
  return d;
}
// Instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n  (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// 8 consecutive-k fp32 values of one operand row -> bf16 hi (and lo) -> one 16-byte k-chunk each
__device__ __forceinline__ void split_store(const float (&v)[8], unsigned char* hi_base, unsigned char* lo_base,
                                            int off, bool x3) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float x0 = v[2 * p], x1 = v[2 * p + 1];
    const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
    h[p] = *reinterpret_cast<const uint32_t*>(&hh);
    const float r0 = x0 - __low2float(hh), r1 = x1 - __high2float(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
    l[p] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (x3) *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Staging items.  An item is (row r of the tile, k-chunk k8 of the 32-wide k-block) = 8 fp32 -> one 16-byte chunk.
//   k-contiguous operands: a warp reads 8 rows x 128 contiguous bytes (two 16-byte loads per thread) and each
//     lane-octet stores one conflict-free 128-byte group;
//   row-contiguous operands (transposed views): id -> (r = id % rows, k8 = id / rows): a warp reads 32
//     consecutive rows per k (coalesced scalar loads) and stores 512 contiguous bytes.
template <class Ld>
__device__ __forceinline__ void item_coords(int id, int rows, int& r, int& k8) {
  // k-contiguous: 8 consecutive lanes take 8 consecutive rows of ONE k-chunk (their 16-byte smem stores fill one
  // conflict-free 128-byte row group); the 4 lane-octets of a warp take the 4 chunks of the same 8 rows, so a warp
  // still reads 8 rows x 128 contiguous bytes from global memory.
  if (Ld::kContig) { r = (id & 7) + ((id >> 5) << 3); k8 = (id >> 3) & 3; } else { r = id % rows; k8 = id / rows; }
}
template <bool FULL, class Ld>
__device__ __forceinline__ void fetch_item(const Ld& ld, const typename Ld::Cur& cur, bool ok, int k, int kend,
                                           float (&v)[8]) {
  if (ok && (FULL || k < kend)) {
    if (FULL) ld.load8_full(cur, v); else ld.load8(cur, kend, v);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
}

struct RegSet { float a[8]; float b0[8]; float b1[8]; };

// FAST = compile-time bf16x3 and operands whose fast_ok() holds (plain rows, every row segment 32-byte aligned: one
// LDG.E.256 per item): the staging code of all full k-blocks carries no alignment state, no tail guards and no precision
// branches; only a final partial k-block runs the guarded path.
template <class LdA, class LdB, class Epi, bool FAST>
__global__ void __launch_bounds__(THREADS, 1) gemm_umma_kernel(LdA a, LdB b, Epi epi, int M, int N, int K, int splitk,
                                                               int BN, int x3_rt) {
  const int x3 = FAST ? 1 : x3_rt;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_empty[2] = {sbase, sbase + 8};        // stage s may be overwritten (tcgen05.commit of its MMAs)
  const uint32_t bar_done = sbase + 16;                     // all MMAs of the tile completed
  const uint32_t bar_full[2] = {sbase + 32, sbase + 40};    // stage s staged by all 16 staging warps
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 24);
  const int b_half_bytes = BN * BK * 2;  // one of hi / lo
  const int stage_bytes = A_STAGE_BYTES + 2 * b_half_bytes;
  unsigned char* stage0 = smem + HEADER_BYTES;

  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int batch = blockIdx.z / splitk, ks = blockIdx.z % splitk;
  int kper = (K + splitk - 1) / splitk;
  kper = ((kper + BK - 1) / BK) * BK;
  const int kbeg = ks * kper;
  const int kend = (kbeg + kper < K) ? kbeg + kper : K;
  const int nkb = kbeg < kend ? (kend - kbeg + BK - 1) / BK : 0;
  const bool tail_full = ((kend - kbeg) & (BK - 1)) == 0;
  const uint32_t ncols = BN <= 32 ? 32u : (BN <= 64 ? 64u : (BN <= 128 ? 128u : 256u));

  if (tid == 0) {
    mbar_init(bar_empty[0], 1);
    mbar_init(bar_empty[1], 1);
    mbar_init(bar_done, 1);
    mbar_init(bar_full[0], STAGE_THREADS / 32);
    mbar_init(bar_full[1], STAGE_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 24, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc(BN);
  const uint32_t a_lbo = BM * 16, b_lbo = BN * 16, sbo = 128;

  // (warp index through __shfl_sync so that the compiler can prove the branch warp-uniform: the descriptors of the
  //  elected issuer thread then live in uniform registers; a `lane == 0` branch wraps every tcgen05.mma in an
  //  elect / R2UR.BROADCAST waterfall loop)
  if (__shfl_sync(0xffffffffu, warp, 0) == STAGE_THREADS / 32) {
    // ================= MMA issuer warp: one thread, decoupled from the staging warps by the full / empty mbarriers
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int st_i = kb & 1;
        mbar_wait(bar_full[st_i], (uint32_t)((kb >> 1) & 1));
        tc_fence_after();
        const uint32_t sa = smem_u32(stage0 + st_i * stage_bytes), sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          const uint32_t a_hi = sa + j * 2 * a_lbo, a_lo = a_hi + BM * BK * 2;
          const uint32_t b_hi = sb + j * 2 * b_lbo, b_lo = b_hi + b_half_bytes;
          const uint64_t dah = make_desc(a_hi, a_lbo, sbo), dbh = make_desc(b_hi, b_lbo, sbo);
          uint32_t acc = (kb > 0 || j > 0) ? 1u : 0u;
          if (x3) {
            const uint64_t dal = make_desc(a_lo, a_lbo, sbo), dbl = make_desc(b_lo, b_lbo, sbo);
            mma_bf16(tmem_base, dal, dbh, idesc, acc);
            mma_bf16(tmem_base, dah, dbl, idesc, 1u);
            acc = 1u;
          }
          mma_bf16(tmem_base, dah, dbh, idesc, acc);
        }
        mma_commit(bar_empty[st_i]);
        if (kb + 1 == nkb) mma_commit(bar_done);
      }
    }
  } else {
  // ================= staging warps (and, afterwards, the epilogue)
  // staging assignment: one A item and up to two B items per thread (see item_coords); one cursor per item
  static_assert(BK == LD_STEP, "loader cursors advance by one k-block");
  int r, k8;
  item_coords<LdA>(tid, BM, r, k8);
  const bool okA = m0 + r < M;
  const int offA = (k8 * BM + r) * 16;
  int kA = kbeg + k8 * 8;
  typename LdA::Cur ca = a.cursor(batch, okA ? m0 + r : 0, kA);
  int offB[2], kB[2];
  bool okB[2];
  typename LdB::Cur cb[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool has = tid + i * STAGE_THREADS < 4 * BN;
    item_coords<LdB>(has ? tid + i * STAGE_THREADS : 0, BN, r, k8);
    okB[i] = has && n0 + r < N;
    offB[i] = has ? (k8 * BN + r) * 16 : -1;
    kB[i] = kbeg + k8 * 8;
    cb[i] = b.cursor(batch, okB[i] ? n0 + r : 0, kB[i]);
  }
  auto fetch = [&](int kb, RegSet& s) {
    if (kb >= nkb) return;
    if (FAST && (kb + 1 < nkb || tail_full)) {
      if (okA) a.load8_fast(ca, s.a); else { _Pragma("unroll") for (int i = 0; i < 8; ++i) s.a[i] = 0.f; }
      if (offB[0] >= 0) { if (okB[0]) b.load8_fast(cb[0], s.b0); else { _Pragma("unroll") for (int i = 0; i < 8; ++i) s.b0[i] = 0.f; } }
      if (offB[1] >= 0) { if (okB[1]) b.load8_fast(cb[1], s.b1); else { _Pragma("unroll") for (int i = 0; i < 8; ++i) s.b1[i] = 0.f; } }
    } else if (kb + 1 < nkb || tail_full) {
      fetch_item<true>(a, ca, okA, kA, kend, s.a);
      if (offB[0] >= 0) fetch_item<true>(b, cb[0], okB[0], kB[0], kend, s.b0);
      if (offB[1] >= 0) fetch_item<true>(b, cb[1], okB[1], kB[1], kend, s.b1);
    } else {
      fetch_item<false>(a, ca, okA, kA, kend, s.a);
      if (offB[0] >= 0) fetch_item<false>(b, cb[0], okB[0], kB[0], kend, s.b0);
      if (offB[1] >= 0) fetch_item<false>(b, cb[1], okB[1], kB[1], kend, s.b1);
    }
    a.advance(ca); kA += BK;
    b.advance(cb[0]); kB[0] += BK;
    b.advance(cb[1]); kB[1] += BK;
  };
  auto stage_and_issue = [&](int kb, const RegSet& s) {
    const int st_i = kb & 1;
    if (kb >= 2) mbar_wait(bar_empty[st_i], (uint32_t)(((kb >> 1) - 1) & 1));  // MMAs that read this stage are done
    unsigned char* st = stage0 + st_i * stage_bytes;
    split_store(s.a, st, st + BM * BK * 2, offA, x3 != 0);
    if (offB[0] >= 0) split_store(s.b0, st + A_STAGE_BYTES, st + A_STAGE_BYTES + b_half_bytes, offB[0], x3 != 0);
    if (offB[1] >= 0) split_store(s.b1, st + A_STAGE_BYTES, st + A_STAGE_BYTES + b_half_bytes, offB[1], x3 != 0);
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive_cta(bar_full[st_i]);  // 16 arrivals (one per staging warp) complete the stage
  };

  // two register sets ping-pong: the loads of k-block kb+1 are in flight while k-block kb is converted and stored
  // (the MMA issue is off this path, so converting one k-block already outlasts the L2 latency of the next)
  RegSet s0, s1;
  fetch(0, s0);
  for (int kb = 0; kb < nkb; kb += 2) {
    fetch(kb + 1, s1);
    stage_and_issue(kb, s0);
    if (kb + 1 < nkb) {
      fetch(kb + 2, s0);
      stage_and_issue(kb + 1, s1);
    }
  }

  }  // staging warps
  if (nkb > 0 && warp < EPI_WARPS) {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    // epilogue: the stage buffers are free now (every MMA has completed); reuse them for the transposition
    float* tbuf = reinterpret_cast<float*>(stage0) + warp * (32 * 33);
    const int lane_base = (warp & 3) * 32;
    for (int c0 = (warp >> 2) * 32; c0 < BN; c0 += 32 * (EPI_WARPS / 4)) {
      uint32_t rr32[32];
      tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c0, rr32);
#pragma unroll
      for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = __uint_as_float(rr32[j]);
      __syncwarp();
      const int n = n0 + c0 + lane;
      const bool n_ok = (c0 + lane < BN) && n < N;
      const int mlim = M - (m0 + lane_base);
      if (n_ok) {
        const typename Epi::Col cc = epi.col(batch, n);  // (batch, n)-dependent state once per column
        const int rmax = mlim < 32 ? mlim : 32;
#pragma unroll 4
        for (int rr = 0; rr < rmax; ++rr) epi.apply(cc, m0 + lane_base + rr, tbuf[rr * 33 + lane], splitk > 1);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

static inline int pick_bn(int N) {
  int tiles = (N + BN_MAX - 1) / BN_MAX;
  int bn = (N + tiles - 1) / tiles;
  bn = ((bn + 15) / 16) * 16;
  if (bn < 16) bn = 16;
  if (bn > BN_MAX) bn = BN_MAX;
  return bn;
}
static inline size_t smem_bytes(int BN) {
  size_t stages = 2 * (size_t)(A_STAGE_BYTES + 2 * BN * BK * 2);
  size_t epi = EPI_WARPS * 32 * 33 * sizeof(float);
  return HEADER_BYTES + (stages > epi ? stages : epi);
}

// splitk > 1 on entry means "the epilogue is linear, partial sums may be combined by atomicAdd": the engine then
// picks its own split so that one wave of CTAs (one per SM) covers the problem.
template <class LdA, class LdB, class Epi, bool FAST>
static inline void launch_variant(const LdA& a, const LdB& b, const Epi& epi, int M, int N, int K, int splitk, int BN,
                                  dim3 grid, void* stream) {
  auto kfn = &gemm_umma_kernel<LdA, LdB, Epi, FAST>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(BN_MAX));
    attr_set = true;
  }
  S2AG_LAUNCH(kfn, grid, THREADS, smem_bytes(BN), stream, a, b, epi, M, N, K, splitk, BN, g_precision == 0 ? 1 : 0);
}

template <class LdA, class LdB, class Epi>
static inline void launch(const LdA& a, const LdB& b, const Epi& epi, int M, int N, int K, int nbatch, int splitk,
                          void* stream) {
  const int BN = pick_bn(N);
  const int tiles = s2ag_cdiv(N, BN) * s2ag_cdiv(M, BM) * nbatch;
  if (splitk > 1) {
    int sk = s2ag_sm_count() / tiles;
    const int maxk = K / (4 * BK);
    if (sk > maxk) sk = maxk;
    if (sk < 1) sk = 1;
    splitk = sk;
  }
  dim3 grid(s2ag_cdiv(N, BN), s2ag_cdiv(M, BM), nbatch * splitk);
  // FAST variant: both operands offer a fast item load, every k-slice is a whole number of k-blocks, fp32-grade mode
  bool fast = false;
  if constexpr (LdA::kHasFast && LdB::kHasFast) {
    int kper = (K + splitk - 1) / splitk;
    kper = ((kper + BK - 1) / BK) * BK;
    (void)kper;
    fast = g_precision == 0 && a.fast_ok() && b.fast_ok() && !(g_dbg_flags & 16);  // (a k tail runs the guarded path)
  }
  if (fast) {
    if constexpr (LdA::kHasFast && LdB::kHasFast) launch_variant<LdA, LdB, Epi, true>(a, b, epi, M, N, K, splitk, BN, grid, stream);
  } else {
    launch_variant<LdA, LdB, Epi, false>(a, b, epi, M, N, K, splitk, BN, grid, stream);
  }
}

// shapes worth a tensor-core tile: anything else stays on the exact-fp32 SIMT kernel
static inline bool worthwhile(int M, int N, int K) { return M >= 8 && N >= 16 && K >= 32; }

}  // namespace umma
}  // namespace s2ag
#endif  // !S2AG_EMU
