// tcgen05 contraction with a PRE-PACKED weight operand:  C[m,n] = epi( sum_k A(m,k) * W(n,k) ).
//
// In gemm_umma.cuh both fp32 operands are split into bf16 hi/lo by the staging warps of every CTA; for a weight
// operand that work is repeated by each of the M/128 row tiles (68x for the 8704-row projections of the hot path) and
// it is ~2/3 of the staging instructions (BN = 232 weight rows vs 128 activation rows per k-block).  Here the weight is
// packed ONCE per call by pack_operand_kernel (into a caller-registered scratch buffer, s2ag_register_scratch) into
// the exact shared-memory image the tensor core reads,
//     [plane hi | lo][k-chunk of 8][row n][16 B]          (K zero-padded to a multiple of 32),
// so that a CTA fetches its B tile with TMA bulk copies (4 chunks x 2 planes per k-block, BN*16 contiguous bytes each,
// complete_tx on the stage's "full" mbarrier) and the staging warps only convert the activation operand (one
// 8-element item per thread and k-block).  Four stages instead of two keep the copies a few k-blocks ahead.
//
// Structure of one CTA (576 threads): 16 staging / epilogue warps, 1 MMA-issuer warp, 1 TMA-producer warp; one
// 128 x BN output tile; barriers  full[s] (16 staging-warp arrivals + 1 expect_tx arrival + the copy bytes),
// empty[s] (tcgen05.commit of the stage's MMAs).  Epilogue as in gemm_umma.cuh.
#pragma once
#include "gemm_umma.cuh"
#ifndef S2AG_EMU

namespace s2ag {
namespace umma {

constexpr int PK_STAGES = 4, PK_THREADS = STAGE_THREADS + 64, PK_HEADER = 256;

// operand image of W: element (n, k) = w[batch*bstride + n*s_n + k*s_k]
struct PackedB {
  const unsigned char* img;  // [batch][plane][chunk][n][16 B]
  long batch_bytes, plane_bytes;
  int nrows, Kpad;
};
static inline long packed_bytes(int N, int K, int nbatch) {
  const long Kpad = (K + BK - 1) / BK * BK;
  return (long)nbatch * 2 * (Kpad / 8) * N * 16;
}

template <class LdB>
__global__ void __launch_bounds__(256) pack_operand_kernel(LdB b, int N, int K, int Kpad, unsigned char* __restrict__ img,
                                                           long batch_bytes, long plane_bytes) {
  const int nchunk = Kpad / 8;
  const long total = (long)nchunk * N;
  const int batch = blockIdx.y;
  unsigned char* ib = img + batch * batch_bytes;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long)gridDim.x * blockDim.x) {
    // k-contiguous operands: adjacent lanes take adjacent 32-byte pieces of one row; row-contiguous ones: adjacent rows
    int n, c;
    if (LdB::kContig) { c = (int)(it % nchunk); n = (int)(it / nchunk); } else { n = (int)(it % N); c = (int)(it / N); }
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (c * 8 + i < K) ? b(batch, n, c * 8 + i) : 0.f;
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * p], v[2 * p + 1]);
      h[p] = *reinterpret_cast<const uint32_t*>(&hh);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * p] - __low2float(hh), v[2 * p + 1] - __high2float(hh));
      l[p] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const long off = ((long)c * N + n) * 16;
    *reinterpret_cast<uint4*>(ib + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(ib + plane_bytes + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// packs the operand seen through loader `b` (nbatch matrices of N rows x K) into `img` (packed_bytes(N, K, nbatch)
// bytes, 16-byte aligned)
template <class LdB>
static inline PackedB pack_operand(const LdB& b, int N, int K, int nbatch, void* img, void* stream) {
  const int Kpad = (K + BK - 1) / BK * BK;
  PackedB pb;
  pb.img = reinterpret_cast<const unsigned char*>(img);
  pb.plane_bytes = (long)(Kpad / 8) * N * 16;
  pb.batch_bytes = 2 * pb.plane_bytes;
  pb.nrows = N; pb.Kpad = Kpad;
  const long total = (long)(Kpad / 8) * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  auto kfn = &pack_operand_kernel<LdB>;
  S2AG_LAUNCH(kfn, dim3(blocks, nbatch), 256, 0, stream, b, N, K, Kpad, reinterpret_cast<unsigned char*>(img),
              pb.batch_bytes, pb.plane_bytes);
  return pb;
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

template <class LdA, class Epi, bool FAST>
__global__ void __launch_bounds__(PK_THREADS, 1) gemm_umma_pk_kernel(LdA a, PackedB pb, Epi epi, int M, int N, int K,
                                                                     int splitk, int BN, int x3_rt) {
  const int x3 = FAST ? 1 : x3_rt;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_empty0 = sbase, bar_full0 = sbase + 64, bar_done = sbase + 128;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 136);
  const int b_half_bytes = BN * BK * 2;  // one of hi / lo
  const int stage_bytes = A_STAGE_BYTES + 2 * b_half_bytes;
  unsigned char* stage0 = smem + PK_HEADER;

  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int batch = blockIdx.z / splitk, ks = blockIdx.z % splitk;
  int kper = (K + splitk - 1) / splitk;
  kper = ((kper + BK - 1) / BK) * BK;
  const int kbeg = ks * kper;
  const int kend = (kbeg + kper < K) ? kbeg + kper : K;
  const int nkb = kbeg < kend ? (kend - kbeg + BK - 1) / BK : 0;
  const bool tail_full = ((kend - kbeg) & (BK - 1)) == 0;
  const uint32_t ncols = BN <= 32 ? 32u : (BN <= 64 ? 64u : (BN <= 128 ? 128u : 256u));
  const int nvalid = (N - n0 < BN) ? N - n0 : BN;   // weight rows of this tile that exist

  if (tid == 0) {
    for (int s = 0; s < PK_STAGES; ++s) {
      mbar_init(bar_empty0 + 8 * s, 1);
      mbar_init(bar_full0 + 8 * s, STAGE_THREADS / 32 + 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 136, ncols);
  if (nvalid < BN) {
    // weight rows beyond N are never copied: clear them once (their accumulator columns are discarded, but must not
    // be fed uninitialised bit patterns that decode to NaN and trap nothing -- kept finite for hygiene)
    for (int s = 0; s < PK_STAGES; ++s) {
      unsigned char* sb = stage0 + s * stage_bytes + A_STAGE_BYTES;
      for (int i = tid; i < 2 * 4 * (BN - nvalid); i += PK_THREADS) {
        const int row = nvalid + i % (BN - nvalid), pc = i / (BN - nvalid);  // pc = plane*4 + chunk
        *reinterpret_cast<uint4*>(sb + (pc >> 2) * b_half_bytes + ((pc & 3) * BN + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_u == STAGE_THREADS / 32) {
    // ================= MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BN);
      const uint32_t a_lbo = BM * 16, b_lbo = BN * 16, sbo = 128;
      for (int kb = 0; kb < nkb; ++kb) {
        const int st_i = kb % PK_STAGES;
        mbar_wait(bar_full0 + 8 * st_i, (uint32_t)((kb / PK_STAGES) & 1));
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
        mma_commit(bar_empty0 + 8 * st_i);
        if (kb + 1 == nkb) mma_commit(bar_done);
      }
    }
  } else if (warp_u == STAGE_THREADS / 32 + 1) {
    // ================= TMA producer: the B tile of k-block kb = chunks (kbeg/8 + 4kb .. +3) x rows [n0, n0 + nvalid)
    if (elect_one()) {
      const unsigned char* gb = pb.img + (long)batch * pb.batch_bytes;
      const uint32_t seg = (uint32_t)nvalid * 16u;
      const uint32_t tx = seg * 4u * (x3 ? 2u : 1u);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st_i = kb % PK_STAGES;
        if (kb >= PK_STAGES) mbar_wait(bar_empty0 + 8 * st_i, (uint32_t)(((kb / PK_STAGES) - 1) & 1));
        const uint32_t bar = bar_full0 + 8 * st_i;
        const uint32_t sb = smem_u32(stage0 + st_i * stage_bytes) + A_STAGE_BYTES;
        mbar_expect_tx(bar, tx);
        const long c0 = (long)(kbeg / 8) + 4L * kb;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const unsigned char* src = gb + ((c0 + c) * pb.nrows + n0) * 16;
          bulk_g2s(sb + (uint32_t)(c * BN * 16), src, seg, bar);
          if (x3) bulk_g2s(sb + (uint32_t)(b_half_bytes + c * BN * 16), src + pb.plane_bytes, seg, bar);
        }
      }
    }
  } else {
    // ================= staging warps: the activation operand, one item per thread and k-block
    static_assert(BK == LD_STEP, "loader cursors advance by one k-block");
    int r, k8;
    item_coords<LdA>(tid, BM, r, k8);
    const bool okA = m0 + r < M;
    const int offA = (k8 * BM + r) * 16;
    int kA = kbeg + k8 * 8;
    typename LdA::Cur ca = a.cursor(batch, okA ? m0 + r : 0, kA);
    auto fetch = [&](int kb, float (&v)[8]) {
      if (kb >= nkb) return;
      if (FAST && (kb + 1 < nkb || tail_full)) {
        if (okA) a.load8_fast(ca, v); else { _Pragma("unroll") for (int i = 0; i < 8; ++i) v[i] = 0.f; }
      } else if (kb + 1 < nkb || tail_full) {
        fetch_item<true>(a, ca, okA, kA, kend, v);
      } else {
        fetch_item<false>(a, ca, okA, kA, kend, v);
      }
      a.advance(ca); kA += BK;
    };
    auto stage_and_arrive = [&](int kb, const float (&v)[8]) {
      const int st_i = kb % PK_STAGES;
      if (kb >= PK_STAGES) mbar_wait(bar_empty0 + 8 * st_i, (uint32_t)(((kb / PK_STAGES) - 1) & 1));
      unsigned char* st = stage0 + st_i * stage_bytes;
      split_store(v, st, st + BM * BK * 2, offA, x3 != 0);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(bar_full0 + 8 * st_i);
    };
    // three register sets in flight: the loads of k-blocks kb+1 and kb+2 are outstanding while kb is converted
    float v0[8], v1[8], v2[8];
    fetch(0, v0);
    fetch(1, v1);
    for (int kb = 0; kb < nkb; kb += 3) {
      fetch(kb + 2, v2);
      stage_and_arrive(kb, v0);
      if (kb + 1 < nkb) {
        fetch(kb + 3, v0);
        stage_and_arrive(kb + 1, v1);
      }
      if (kb + 2 < nkb) {
        fetch(kb + 4, v1);
        stage_and_arrive(kb + 2, v2);
      }
    }
  }
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
        const typename Epi::Col cc = epi.col(batch, n);
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

static inline size_t pk_smem_bytes(int BN) {
  size_t stages = PK_STAGES * (size_t)(A_STAGE_BYTES + 2 * BN * BK * 2);
  size_t epi = EPI_WARPS * 32 * 33 * sizeof(float);
  return PK_HEADER + (stages > epi ? stages : epi);
}

template <class LdA, class Epi, bool FAST>
static inline void pk_launch_variant(const LdA& a, const PackedB& pb, const Epi& epi, int M, int N, int K, int splitk,
                                     int BN, dim3 grid, void* stream) {
  auto kfn = &gemm_umma_pk_kernel<LdA, Epi, FAST>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pk_smem_bytes(BN_MAX));
    attr_set = true;
  }
  S2AG_LAUNCH(kfn, grid, PK_THREADS, pk_smem_bytes(BN), stream, a, pb, epi, M, N, K, splitk, BN, g_precision == 0 ? 1 : 0);
}

// C = epi(A * W^T) with W given as a packed image (pack_operand); same split-K convention as umma::launch
template <class LdA, class Epi>
static inline void launch_packed(const LdA& a, const PackedB& pb, const Epi& epi, int M, int N, int K, int nbatch,
                                 int splitk, void* stream) {
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
  bool fast = false;
  if constexpr (LdA::kHasFast) fast = g_precision == 0 && a.fast_ok() && !(g_dbg_flags & 16);
  if (fast) {
    if constexpr (LdA::kHasFast) pk_launch_variant<LdA, Epi, true>(a, pb, epi, M, N, K, splitk, BN, grid, stream);
  } else {
    pk_launch_variant<LdA, Epi, false>(a, pb, epi, M, N, K, splitk, BN, grid, stream);
  }
}

}  // namespace umma
}  // namespace s2ag
#endif  // !S2AG_EMU
