// Stride-1 Conv1d / Conv2d over channels-last activations as a SHIFTED-WINDOW contraction on tcgen05 (sm_100a).
// Replaces the implicit-im2col GEMM for the small-channel convolutions of AffEncoder / STGraphConv / MFCCEncoder /
// ConvDiscriminator (net/multimodal_context_net_v2.py:39-45,:146-150,:397-404; net/utils/tgcn.py:64-68,181-187,199-203)
// whose weights fit in shared memory, forward and data-gradient.
//
// im2col re-reads every input pixel KH*KW times (45x for the 9x5 ST-GCN kernel).  Here the zero-padded image is
// linearised, q = (n*Hp + hp)*Wp + wp, so that filter tap (kh, kw) is the CONSTANT row shift kh*Wp + kw:
//   * a CTA stages the padded pixels [a0, a0 + 128 + (KH-1)*Wp + KW-1) of its 128-anchor tile ONCE into shared memory as
//     bf16 hi/lo in the K-major UMMA layout [channel chunk of 8][pixel row][16 B];
//   * tap t is then one tcgen05.mma per 16 input channels whose A descriptor simply starts `shift(t)` rows further down
//     the same image; B = that tap's [Cout x Cin] weight slab, stationary in shared memory for the CTA's lifetime
//     (persistent CTAs loop over tiles);
//   * four issuer threads (elected lanes of warp-uniform branches) take the taps round-robin into four TMEM
//     accumulators that the epilogue sums; anchors in the padding (hp >= Ho or wp >= Wo) are discarded.
// Data-gradient = the same kernel over dY with flipped taps, padding (K-1-p) and the transposed weight view.
#include "s2ag.h"
#include "gemm_umma.cuh"

namespace s2ag {
namespace convs {

using namespace s2ag::umma;

constexpr int CBM = 128, CTHREADS = 384, CWORK = 256, CHDR = 256, NISS = 4;

struct Params {
  const float* x; long ldpix_x; int N, H, W, Cin;   // source image (forward: x; data-gradient: dY), channels-last
  const float* w; int w_mode, w_ci, KK;             // reference weight [.][w_ci][KK]; mode 0 forward, 1 data-gradient
  const float* bias; float* y; long ldpix_y; int Cout;
  int KH, KW, ph, pw;                               // anchor padding
  int Ho, Wo, Hp, Wp;
  int act; float slope; int accumulate;
  int Kc, Np, R, tiles; long total_rows;            // channel chunks, padded Cout, staged rows per tile, tiles, anchors
  int x3;
};

__device__ __forceinline__ void pack8c(const float (&v)[8], uint4& hi, uint4& lo) {
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
__device__ __forceinline__ void tmem_ld8c(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}

__global__ void __launch_bounds__(CTHREADS, 2) conv_shift_kernel(Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // provably warp-uniform: issuer descriptors in uniform registers
  const uint32_t sbase = smem_u32(smem);
  const uint32_t mma_bar = sbase;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 16);
  const int Kc = p.Kc, Np = p.Np, R = p.R, KK = p.KK;
  const int w_half = KK * Kc * Np * 16;
  unsigned char* w_hi = smem + CHDR;                  // [tap][k-chunk][Np][16 B]
  unsigned char* w_lo = w_hi + w_half;
  const int a_half = Kc * R * 16;
  unsigned char* a_hi = w_lo + w_half;                // [k-chunk][R][16 B]
  unsigned char* a_lo = a_hi + a_half;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(NISS * Np)) ncols <<= 1;

  if (tid == 0) {
    mbar_init(mma_bar, NISS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 16, ncols);

  // ---- stationary weights: element (tap, n, k) of the selected view, zero padded to [Np][Kc*8]
  for (int idx = tid; idx < KK * Kc * Np; idx += CTHREADS) {
    const int n = idx % Np; const int kc = (idx / Np) % Kc; const int t = idx / (Np * Kc);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kc * 8 + i;
      float val = 0.f;
      if (n < p.Cout && k < p.Cin) {
        val = p.w_mode == 0 ? __ldg(p.w + ((long)n * p.w_ci + k) * KK + t)
                            : __ldg(p.w + ((long)k * p.w_ci + n) * KK + (KK - 1 - t));
      }
      v[i] = val;
    }
    uint4 hi, lo;
    pack8c(v, hi, lo);
    *reinterpret_cast<uint4*>(w_hi + idx * 16) = hi;
    *reinterpret_cast<uint4*>(w_lo + idx * 16) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // accumulator i is written iff issuer i has at least one tap
  const int n_acc = KK < NISS ? KK : NISS;
  const bool vec_src = (p.ldpix_x & 3) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
  const int HpWp = p.Hp * p.Wp;
  const bool vec_dst = (p.ldpix_y & 7) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0;

  uint32_t parity = 0;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, parity ^= 1u) {
    const long a0 = (long)tile * CBM;
    if (warp < 8) {
      // ---- stage the padded image rows [a0, a0 + R) once: item = (k-chunk, row)
      for (int it = tid; it < Kc * R; it += CWORK) {
        const int r = it % R, kc = it / R;
        const long q = a0 + r;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (q < p.total_rows) {
          const int n = (int)(q / HpWp); const int rem = (int)(q - (long)n * HpWp);
          const int hs = rem / p.Wp - p.ph, ws = rem % p.Wp - p.pw;
          if (hs >= 0 && hs < p.H && ws >= 0 && ws < p.W) {
            const float* src = p.x + ((long)(n * p.H + hs) * p.W + ws) * p.ldpix_x + kc * 8;
            if (vec_src && kc * 8 + 8 <= p.Cin) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
              v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (kc * 8 + i < p.Cin) v[i] = __ldg(src + i);
            }
          }
        }
        uint4 hi, lo;
        pack8c(v, hi, lo);
        *reinterpret_cast<uint4*>(a_hi + it * 16) = hi;
        if (p.x3) *reinterpret_cast<uint4*>(a_lo + it * 16) = lo;
      }
      fence_async_smem();
    }
    __syncthreads();
    if (warp_u >= 8 && elect_one()) {
      // ---- issuer `iss`: taps iss, iss + NISS, ... into accumulator iss
      const int iss = warp_u - 8;
      tc_fence_after();
      const uint32_t idesc = make_idesc(Np);
      const uint32_t a_lbo = (uint32_t)R * 16, w_lbo = (uint32_t)Np * 16;
      const uint32_t sa = smem_u32(a_hi), sw = smem_u32(w_hi);
      const uint32_t d = tmem_base + (uint32_t)(iss * Np);
      uint32_t cnt = 0;
      for (int t = iss; t < KK; t += NISS) {
        const int shift = (t / p.KW) * p.Wp + (t % p.KW);
        for (int k2 = 0; k2 < (Kc >> 1); ++k2) {
          const uint32_t ah = sa + (uint32_t)((2 * k2 * R + shift) * 16), al = ah + (uint32_t)a_half;
          const uint32_t wh = sw + (uint32_t)(((t * Kc + 2 * k2) * Np) * 16), wl = wh + (uint32_t)w_half;
          const uint64_t dah = make_desc(ah, a_lbo, 128), dwh = make_desc(wh, w_lbo, 128);
          if (p.x3) {
            mma_bf16(d, make_desc(al, a_lbo, 128), dwh, idesc, cnt ? 1u : 0u); ++cnt;
            mma_bf16(d, dah, make_desc(wl, w_lbo, 128), idesc, 1u); ++cnt;
          }
          mma_bf16(d, dah, dwh, idesc, cnt ? 1u : 0u); ++cnt;
        }
      }
      mma_commit(mma_bar);
    }
    mbar_wait(mma_bar, parity);
    tc_fence_after();
    if (warp < 8) {
      // ---- epilogue: thread = anchor row, warpgroup = half of the output channels
      const int row = (warp & 3) * 32 + lane;
      const long q = a0 + row;
      const int n = (int)(q / HpWp); const int rem = (int)(q - (long)n * HpWp);
      const int ho = rem / p.Wp, wo = rem % p.Wp;
      const bool ok = q < p.total_rows && ho < p.Ho && wo < p.Wo;
      float* dst = p.y + ((long)(n * p.Ho + ho) * p.Wo + wo) * p.ldpix_y;
      const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      const int c_beg = (warp >> 2) * (Np >> 1), c_end = c_beg + (Np >> 1);
      for (int c0 = c_beg; c0 < c_end; c0 += 8) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int a = 0; a < NISS; ++a) {
          if (a < n_acc) {
            float v[8];
            tmem_ld8c(t_lane + (uint32_t)(a * Np + c0), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v[i];
          }
        }
        if (ok) {
          if (vec_dst && c0 + 8 <= p.Cout) {
            // 8 consecutive channels of one pixel = one 32-byte sector
            float val[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              val[i] = acc[i] + (p.bias ? __ldg(p.bias + c0 + i) : 0.f);
              val[i] = s2ag_act(val[i], p.act, p.slope);
            }
            if (p.accumulate) {
              const float4 o0 = *reinterpret_cast<const float4*>(dst + c0), o1 = *reinterpret_cast<const float4*>(dst + c0 + 4);
              val[0] += o0.x; val[1] += o0.y; val[2] += o0.z; val[3] += o0.w;
              val[4] += o1.x; val[5] += o1.y; val[6] += o1.z; val[7] += o1.w;
            }
            asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"f"(val[0]), "f"(val[1]), "f"(val[2]),
                         "f"(val[3]), "f"(val[4]), "f"(val[5]), "f"(val[6]), "f"(val[7]), "l"(dst + c0)
                         : "memory");
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = c0 + i;
              if (c < p.Cout) {
                float val = acc[i];
                if (p.bias) val += __ldg(p.bias + c);
                val = s2ag_act(val, p.act, p.slope);
                if (p.accumulate) dst[c] += val; else dst[c] = val;
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

}  // namespace convs

static inline int cs_round_up(int a, int b) { return (a + b - 1) / b * b; }

// true (and launches) when the shifted-window kernel applies; false -> the caller uses the implicit-GEMM engine
bool conv_shift_launch(const float* x, long ldpix_x, int N, int H, int W, int Cin, const float* w, int w_mode, int w_ci,
                       const float* bias, float* y, long ldpix_y, int Cout, int KH, int KW, int ph, int pw, int Ho,
                       int Wo, int act, float slope, int accumulate, void* stream) {
  using namespace convs;
  if (umma::g_dbg_flags & 4) return false;  // bring-up switch: force the implicit-GEMM path
  const int Cin_pad = cs_round_up(Cin, 16), Np = cs_round_up(Cout, 16);
  if (Np > 512 / NISS || ph < 0 || pw < 0) return false;
  const int Hp = H + 2 * ph, Wp = W + 2 * pw;
  if (Hp < KH || Wp < KW) return false;
  const int KK = KH * KW, Kc = Cin_pad / 8;
  const int R = cs_round_up(CBM + (KH - 1) * Wp + (KW - 1), 8);
  const size_t smem = CHDR + 2 * (size_t)KK * Kc * Np * 16 + 2 * (size_t)Kc * R * 16;
  if (smem > 220 * 1024 || R > 8192) return false;
  const long total = (long)N * Hp * Wp;
  if (total <= 0 || total > (1L << 30)) return false;
  // work heuristic: worthwhile only when the im2col redundancy is real or the GEMM would be skinny
  Params p;
  p.x = x; p.ldpix_x = ldpix_x; p.N = N; p.H = H; p.W = W; p.Cin = Cin;
  p.w = w; p.w_mode = w_mode; p.w_ci = w_ci; p.KK = KK; p.bias = bias; p.y = y; p.ldpix_y = ldpix_y; p.Cout = Cout;
  p.KH = KH; p.KW = KW; p.ph = ph; p.pw = pw; p.Ho = Ho; p.Wo = Wo; p.Hp = Hp; p.Wp = Wp;
  p.act = act; p.slope = slope; p.accumulate = accumulate;
  p.Kc = Kc; p.Np = Np; p.R = R; p.total_rows = total; p.tiles = (int)((total + CBM - 1) / CBM);
  p.x3 = umma::g_precision == 0 ? 1 : 0;
  auto kfn = &conv_shift_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(NISS * Np)) ncols <<= 1;
  const int by_tmem = 512 / (int)ncols;
  if (per_sm > by_tmem) per_sm = by_tmem;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;  // register budget (__launch_bounds__(384, 2))
  // keep the hardware from co-scheduling more CTAs per SM than TMEM columns allow (a blocked tcgen05.alloc would
  // serialise persistent CTAs): pad the dynamic shared memory request accordingly
  size_t smem_req = smem;
  const size_t min_req = (size_t)(225 * 1024) / (per_sm + 1) + 2048;
  if (smem_req < min_req) smem_req = min_req;
  int grid = s2ag_sm_count() * per_sm;
  if (grid > p.tiles) grid = p.tiles;
  S2AG_LAUNCH(kfn, grid, CTHREADS, smem_req, stream, p);
  return true;
}

}  // namespace s2ag
