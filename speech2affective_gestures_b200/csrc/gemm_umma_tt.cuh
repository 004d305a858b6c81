// tcgen05 contraction with BOTH operands pre-packed and fetched by TMA ("tt"): C[m,n] = epi(sum_k A(m,k) B(n,k)).
//
// gemm_umma_packed.cuh still converts the activation operand fp32 -> bf16 hi/lo inside the main loop (16 staging warps;
// tensor pipe 40 % active, the kernel is bound by that conversion: profiles/r01_ncu_gemm_umma_pk.txt).  Here the
// conversion leaves the main loop altogether:
//   * pack_rows_kernel<Ld>: one pass over the operand through its loader cursor (vector loads; the implicit im2col of a
//     convolution is materialised here, once, instead of once per column tile) writes the bf16 hi/lo operand image
//     [plane][8-wide k-chunk][row][16 B], rows padded with zeros to the tile multiple;
//   * gemm_umma_tt_kernel<Epi>: one CTA = one 256 x BN output tile = TWO 128-row accumulators in TMEM sharing every B
//     tile (half the weight traffic per output of the 128-row kernels); a TMA warp streams the k-blocks of both images
//     through a 3-stage ring (cp.async.bulk, 62 KB per stage at BN = 232), THREE issuer threads take the k-blocks in
//     turn (an issuing thread blocks ~100 cycles per tcgen05.mma and ~180 cycles per mbarrier wait, measured in
//     umma_tcn.cu: one thread cannot keep the pipe busy), 16 epilogue warps run the same transposing epilogue as the
//     other engines once both accumulators are complete.
// MEASURED (tools/tt_timeline.cu, GRU projection 8704 x 1800 x 600, bf16x3): main loop 30.8 k cycles per 256 x 232 tile =
// 1.6 k per k-block against 1.26 k of MMA time (the three issuers keep the pipe 80 % busy; one issuer: 40 %), but then a
// 23 k-cycle epilogue: the 148 CTAs of a wave finish together and write 35 MB at once, ~3 TB/s of L2 write bandwidth,
// whatever the store pattern (row-vector and transposing epilogues cost the same).  With the 17 us of packing passes:
// 86 us against 93 us for the packed-B kernel, and no difference inside the GAN step (13.75 vs 13.68 ms).  The missing
// piece is overlapping the epilogue with the next tile's main loop, which needs a second accumulator buffer: 2 x 2 x 232
// TMEM columns do not exist, i.e. 128-row CTAs -- whose operand ingest (72 B/clk for the MMAs to stay busy) exceeds what
// one SM pulls from L2 -- unless a CTA PAIR shares the B tile (cta_group::2 with the peer's copy signalling the leader
// directly: umma_tcn.cu found the relay hop to be the obstacle).  OPT-IN (s2ag_debug_flags bit 16384), parity-tested.
// Routed (gemm.cuh) for the big weight-side contractions: M >= 1024 rows, a packable A loader, scratch for both images.
#pragma once
#include "gemm_umma_packed.cuh"
#ifndef S2AG_EMU

namespace s2ag {
namespace umma {

constexpr int TT_BM = 256, TT_STAGES = 3, TT_NISS = 3, TT_EPI_WARPS = 16;
constexpr int TT_THREADS = (TT_EPI_WARPS + 1 + TT_NISS) * 32, TT_HEADER = 256;

// which loaders may be packed as the A operand, and whether their rows depend on the batch index
template <class L> struct TtTraits { static constexpr bool kPackable = false; static bool batch_invariant(const L&) { return false; } };
template <> struct TtTraits<LdPlain<true>> {
  static constexpr bool kPackable = true;
  static bool batch_invariant(const LdPlain<true>& l) { return l.bstride == 0; }
};
template <> struct TtTraits<LdConv<ORDER_KKC>> {
  static constexpr bool kPackable = true;
  static bool batch_invariant(const LdConv<ORDER_KKC>&) { return true; }
};

// operand image of `rows` x K through loader `ld`, TILED so that the operand of one (row tile, k-block) is ONE contiguous
// piece = one bulk copy (the copy engine costs ~240 cycles per cp.async.bulk whatever its size: 16 copies of 4 KB per
// k-block made the kernel copy-issue-bound, measured):  [batch][row tile][k-block][plane hi|lo][chunk 4][tile_rows][16 B];
// rows >= rows and k >= K are zero
template <class Ld>
__global__ void __launch_bounds__(256) pack_rows_kernel(Ld ld, int rows, int rows_pad, int tile_rows, int K, int Kpad,
                                                        unsigned char* __restrict__ img, long batch_bytes) {
  const int nchunk = Kpad / 8, nkb = Kpad / BK;
  const long total = (long)nchunk * rows_pad;
  const int batch = blockIdx.y;
  unsigned char* ib = img + batch * batch_bytes;
  const long plane = 4L * tile_rows * 16;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long)gridDim.x * blockDim.x) {
    // k-contiguous operands: adjacent lanes take adjacent 32-byte pieces of one row; row-contiguous ones: adjacent rows
    int r, c;
    if (Ld::kContig) { c = (int)(it % nchunk); r = (int)(it / nchunk); } else { r = (int)(it % rows_pad); c = (int)(it / rows_pad); }
    float v[8];
    if (r < rows && c * 8 < K) {
      const typename Ld::Cur cu = ld.cursor(batch, r, c * 8);
      ld.load8(cu, K, v);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * p], v[2 * p + 1]);
      h[p] = *reinterpret_cast<const uint32_t*>(&hh);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * p] - __low2float(hh), v[2 * p + 1] - __high2float(hh));
      l[p] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const int t = r / tile_rows, rr = r - t * tile_rows, kb = c >> 2, cc = c & 3;
    const long off = ((long)t * nkb + kb) * 2 * plane + ((long)cc * tile_rows + rr) * 16;
    *reinterpret_cast<uint4*>(ib + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(ib + off + plane) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

struct PackedRows {
  const unsigned char* img;
  long batch_bytes;                // 0: one image shared by every batch
  int tile_rows, nkb;              // rows per tile, k-blocks of the whole (padded) K
};
static inline long packed_rows_bytes(int rows_pad, int K, int nbatch) {
  const long Kpad = (K + BK - 1) / BK * BK;
  return (long)nbatch * 2 * (Kpad / 8) * rows_pad * 16;
}
template <class Ld>
static inline PackedRows pack_rows(const Ld& ld, int rows, int rows_pad, int tile_rows, int K, int nbatch, void* img,
                                   void* stream) {
  const int Kpad = (K + BK - 1) / BK * BK;
  PackedRows pr;
  pr.img = reinterpret_cast<const unsigned char*>(img);
  pr.batch_bytes = 2L * (Kpad / 8) * rows_pad * 16;
  pr.tile_rows = tile_rows; pr.nkb = Kpad / BK;
  const long total = (long)(Kpad / 8) * rows_pad;
  long blocks = (total + 255) / 256;
  const long cap = (long)s2ag_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  auto kfn = &pack_rows_kernel<Ld>;
  S2AG_LAUNCH(kfn, dim3((unsigned)blocks, nbatch), 256, 0, stream, ld, rows, rows_pad, tile_rows, K, Kpad,
              reinterpret_cast<unsigned char*>(img), pr.batch_bytes);
  return pr;
}

#ifdef S2AG_TT_TIMELINE
__device__ long long g_tt_tl[4][32];   // [producer | issuer 0 waited | issuer 0 issued | misc][k-block]
#define TT_MARK(r, i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (i) < 32) g_tt_tl[r][i] = clock64(); } while (0)
#else
#define TT_MARK(r, i) do { } while (0)
#endif

template <class Epi>
__global__ void __launch_bounds__(TT_THREADS, 1) gemm_umma_tt_kernel(PackedRows pa, PackedRows pb, Epi epi, int M, int N, int K,
                                                                     int splitk, int BN, int x3) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full0 = sbase, bar_empty0 = sbase + 32, bar_done = sbase + 64, bar_ord = sbase + 72;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 80);
  const int a_plane = 4 * TT_BM * 16;              // one plane of one stage: 4 chunks x 256 rows x 16 B
  const int b_plane = 4 * BN * 16;
  const int stage_bytes = 2 * a_plane + 2 * b_plane;
  unsigned char* stage0 = smem + TT_HEADER;

  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TT_BM;
  const int batch = blockIdx.z / splitk, ks = blockIdx.z % splitk;
  int kper = (K + splitk - 1) / splitk;
  kper = ((kper + BK - 1) / BK) * BK;
  const int kbeg = ks * kper;
  const int kend = (kbeg + kper < K) ? kbeg + kper : K;
  const int nkb = kbeg < kend ? (kend - kbeg + BK - 1) / BK : 0;
  const bool two = m0 + BM < M;                    // the second 128-row accumulator has valid rows
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(2 * BN)) ncols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < TT_STAGES; ++s) { mbar_init(bar_full0 + 8 * s, 1); mbar_init(bar_empty0 + 8 * s, 1); }
    mbar_init(bar_done, TT_NISS);
    mbar_init(bar_ord, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 80, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) TT_MARK(3, 0);

  if (warp_u == TT_EPI_WARPS) {
    // ================= TMA producer: k-block kb = chunks (kbeg/8 + 4 kb .. +3) of both images
    if (elect_one()) {
      const unsigned char* ga = pa.img + (long)batch * pa.batch_bytes;
      const unsigned char* gb = pb.img + (long)batch * pb.batch_bytes;
      const uint32_t abytes = (uint32_t)(a_plane * (x3 ? 2 : 1)), bbytes = (uint32_t)(b_plane * (x3 ? 2 : 1));
      const long kb0 = kbeg / BK;
      const unsigned char* srca = ga + ((long)blockIdx.y * pa.nkb + kb0) * (2L * a_plane);
      const unsigned char* srcb = gb + ((long)blockIdx.x * pb.nkb + kb0) * (2L * b_plane);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st_i = kb % TT_STAGES;
        if (kb >= TT_STAGES) mbar_wait(bar_empty0 + 8 * st_i, (uint32_t)(((kb / TT_STAGES) - 1) & 1));
        const uint32_t bar = bar_full0 + 8 * st_i;
        const uint32_t sa = smem_u32(stage0 + st_i * stage_bytes), sb = sa + 2 * a_plane;
        TT_MARK(0, kb);
        mbar_expect_tx(bar, abytes + bbytes);
        bulk_g2s(sa, srca + (long)kb * (2L * a_plane), abytes, bar);     // hi (and lo) planes of the A tile: one copy
        bulk_g2s(sb, srcb + (long)kb * (2L * b_plane), bbytes, bar);
      }
    }
  } else if (warp_u > TT_EPI_WARPS) {
    // ================= MMA issuers: issuer i takes the k-blocks kb = i, i + TT_NISS, ...
    const int iss = warp_u - TT_EPI_WARPS - 1;
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BN);
      const uint32_t a_lbo = TT_BM * 16, b_lbo = BN * 16, sbo = 128;
      // the accumulators are initialised by k-block 0 (issuer 0): the others start once those MMAs are issued
      if (iss > 0 && nkb > iss) mbar_wait(bar_ord, 0);
      for (int kb = iss; kb < nkb; kb += TT_NISS) {
        const int st_i = kb % TT_STAGES;
        mbar_wait(bar_full0 + 8 * st_i, (uint32_t)((kb / TT_STAGES) & 1));
        TT_MARK(1, kb);
        tc_fence_after();
        const uint32_t sa = smem_u32(stage0 + st_i * stage_bytes), sb = sa + 2 * a_plane;
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
          if (rt == 1 && !two) break;
          const uint32_t d = tmem_base + (uint32_t)(rt * BN);
#pragma unroll
          for (int j = 0; j < BK / 16; ++j) {
            const uint32_t a_hi = sa + (uint32_t)((j * 2 * TT_BM + rt * BM) * 16), a_lo = a_hi + a_plane;
            const uint32_t b_hi = sb + (uint32_t)(j * 2 * b_lbo), b_lo = b_hi + b_plane;
            const uint64_t dah = make_desc(a_hi, a_lbo, sbo), dbh = make_desc(b_hi, b_lbo, sbo);
            uint32_t acc = (kb > 0 || j > 0) ? 1u : 0u;
            if (x3) {
              mma_bf16(d, make_desc(a_lo, a_lbo, sbo), dbh, idesc, acc);
              mma_bf16(d, dah, make_desc(b_lo, b_lbo, sbo), idesc, 1u);
              acc = 1u;
            }
            mma_bf16(d, dah, dbh, idesc, acc);
          }
        }
        if (kb == 0) mbar_arrive_cta(bar_ord);
        mma_commit(bar_empty0 + 8 * st_i);
        TT_MARK(2, kb);
      }
      mma_commit(bar_done);   // every issuer: all of its MMAs have completed
    }
  }
  if (nkb > 0 && warp < TT_EPI_WARPS) {
    // ================= epilogue (both accumulators complete; the stage buffers are free: reuse them for the transposition)
    {
      uint32_t spins = 0;
      while (!mbar_try_wait(bar_done, 0)) { __nanosleep(256); if (++spins > (1u << 22)) __trap(); }
    }
    tc_fence_after();
    if (tid == 0) TT_MARK(3, 1);
    const int lane_base = (warp & 3) * 32;
    bool done_vec = false;
    if constexpr (EpiHasRowVec<Epi>::value) {
      if (epi.row_vec_ok(splitk > 1) && (n0 & 3) == 0) {
        // ---- row-vector epilogue: thread = output row, a quarter of the tile's 8-column groups per warp of a lane
        //      quadrant; two 16-byte stores per group, no transposition (the transposing walk below spent 27 k cycles
        //      per 256 x 232 tile, as long as the main loop: tools/tt_timeline.cu)
        const float* bsrc = epi.bias_of(batch);
        const int groups = BN / 8, per = (groups + 3) / 4, cq = warp >> 2;
        const int g_beg = cq * per, g_end = g_beg + per < groups ? g_beg + per : groups;
        for (int rt = 0; rt < (two ? 2 : 1); ++rt) {
          const int m = m0 + rt * BM + lane_base + lane;
          const uint32_t t_row = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(rt * BN);
          for (int g = g_beg; g < g_end; ++g) {
            uint32_t u[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                         : "r"(t_row + (uint32_t)(g * 8)) : "memory");
            const int n = n0 + g * 8;
            float b8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) b8[i] = (bsrc && n + i < N) ? __ldg(bsrc + n + i) : 0.f;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < M) {
              float acc[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[i] = __uint_as_float(u[i]);
              if (n + 8 <= N) {
                epi.store_row8(batch, m, n, acc, b8);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (n + i < N) epi(batch, m, n + i, acc[i], false);
              }
            }
          }
        }
        done_vec = true;
      }
    }
    float* tbuf = reinterpret_cast<float*>(stage0) + warp * (32 * 33);
    for (int rt = 0; rt < (done_vec ? 0 : (two ? 2 : 1)); ++rt) {
      const int mrow0 = m0 + rt * BM + lane_base;
      for (int c0 = (warp >> 2) * 32; c0 < BN; c0 += 32 * (TT_EPI_WARPS / 4)) {
        uint32_t rr32[32];
        tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(rt * BN + c0), rr32);
#pragma unroll
        for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = __uint_as_float(rr32[j]);
        __syncwarp();
        const int n = n0 + c0 + lane;
        const bool n_ok = (c0 + lane < BN) && n < N;
        const int mlim = M - mrow0;
        if (n_ok && mlim > 0) {
          const typename Epi::Col cc = epi.col(batch, n);
          const int rmax = mlim < 32 ? mlim : 32;
#pragma unroll 4
          for (int rr = 0; rr < rmax; ++rr) epi.apply(cc, mrow0 + rr, tbuf[rr * 33 + lane], splitk > 1);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) TT_MARK(3, 2);
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

static inline size_t tt_smem_bytes(int BN) {
  const size_t stages = TT_STAGES * (size_t)(2 * 4 * TT_BM * 16 + 2 * 4 * BN * 16);
  const size_t epi = TT_EPI_WARPS * 32 * 33 * sizeof(float);
  return TT_HEADER + (stages > epi ? stages : epi);
}

// true (and launched) when the shape / loaders / scratch allow the two-TMA kernel; false: the caller takes another route
template <class LdA, class LdB, class Epi>
static inline bool launch_tt(const LdA& a, const LdB& b, const Epi& epi, int M, int N, int K, int nbatch, int splitk,
                             void* (*scratch)(void*, long), void* stream) {
  if constexpr (!TtTraits<LdA>::kPackable) {
    return false;
  } else {
    if (M < 4 * TT_BM || K < 2 * BK || !(g_dbg_flags & 16384)) return false;   // opt-in: see the header comment
    const int BN = pick_bn(N);
    const int tilesN = s2ag_cdiv(N, BN), tilesM = s2ag_cdiv(M, TT_BM);
    // 256-row tiles halve the CTA count: only where they still fill the device (else the 128-row packed-B kernel)
    if ((long)tilesN * tilesM * nbatch * (splitk > 1 ? 1 : 1) < (long)s2ag_sm_count() * 9 / 10) return false;
    const int rowsA = tilesM * TT_BM, rowsB = tilesN * BN;
    const int a_batches = TtTraits<LdA>::batch_invariant(a) ? 1 : nbatch;
    const long bytesB = packed_rows_bytes(rowsB, K, nbatch), bytesA = packed_rows_bytes(rowsA, K, a_batches);
    unsigned char* img = reinterpret_cast<unsigned char*>(scratch(stream, bytesA + bytesB + 256));
    if (img == nullptr) return false;
    const int tiles = tilesN * tilesM * nbatch;
    if (splitk > 1) {
      int sk = s2ag_sm_count() / tiles;
      const int maxk = K / (4 * BK);
      if (sk > maxk) sk = maxk;
      if (sk < 1) sk = 1;
      splitk = sk;
    }
    PackedRows pb = pack_rows(b, N, rowsB, BN, K, nbatch, img, stream);
    unsigned char* imgA = img + ((bytesB + 255) & ~255L);
    PackedRows pa = pack_rows(a, M, rowsA, TT_BM, K, a_batches, imgA, stream);
    if (a_batches == 1) pa.batch_bytes = 0;
    auto kfn = &gemm_umma_tt_kernel<Epi>;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tt_smem_bytes(BN_MAX));
      attr_set = true;
    }
    dim3 grid(tilesN, tilesM, nbatch * splitk);
    S2AG_LAUNCH(kfn, grid, TT_THREADS, tt_smem_bytes(BN), stream, pa, pb, epi, M, N, K, splitk, BN, g_precision == 0 ? 1 : 0);
    return true;
  }
}

}  // namespace umma
}  // namespace s2ag
#endif  // !S2AG_EMU
