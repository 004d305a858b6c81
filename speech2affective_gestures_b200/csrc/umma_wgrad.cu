// Weight gradient of the stride-1 Conv1d / Conv2d layers over channels-last activations as a SHIFTED-WINDOW
// contraction over pixels on tcgen05 (sm_100a).  Companion of umma_conv.cu (forward / data-gradient) for the
// small-channel convolutions of AffEncoder / STGraphConv / MFCCEncoder / ConvDiscriminator
// (net/multimodal_context_net_v2.py:39-45,:146-150,:397-404; net/utils/tgcn.py:64-68,181-187,199-203), whose
// backward pass torch.autograd runs through cudnn_convolution_backward_weight in the reference.
//
//   dW[co][ci][tap] += sum over output pixels  dY[pix][co] * X[pix + shift(tap)][ci]
//
// The implicit-GEMM formulation (M = Cout, N = Cin*KH*KW, K = pixels) gathers every activation KH*KW times, element
// by element, for a 128-row tile that is mostly padding (Cout = 16..80).  Here the zero-padded image is linearised
// exactly as in umma_conv.cu, q = (n*Hp + hp)*Wp + wp, and a CTA stages, per tile of 128 anchors,
//   * the padded X rows [a0, a0 + 128 + max shift) and
//   * the dY rows of the 128 anchors (zero where the anchor lies in the padding)
// ONCE into shared memory as bf16 hi/lo images [channel group of 8][pixel row][16 B].  Read as an MN-MAJOR tcgen05
// operand (MN = channel, K = pixel) this layout is linear in the pixel index (k-block stride 128 B = 8 rows), so
// filter tap t is the SAME image with its descriptor start address moved `shift(t)` rows down: one tcgen05.mma per
// 16 pixels and tap, no per-tap copies.  The KH*KW accumulators [channel x channel] stay in TMEM across all tiles of
// a (persistent) CTA and are added to dW with red.global once at the end.
// The wider channel count (<= 128) takes the M side, the other one the N side (KH*KW * pad16(N side) <= 512 columns).
#include "s2ag.h"
#include "gemm_umma.cuh"

namespace s2ag {
namespace wgrads {

using namespace s2ag::umma;

constexpr int WBM = 128;       // anchors (= contraction length) per tile
constexpr int WTHREADS = 384;  // 8 staging / epilogue warps + 4 MMA-issuer warps (8 issuers: no gain, measured)
constexpr int WWORK = 256, WHDR = 256, WISS = 4;

struct Params {
  const float* x; long ldpix_x; int N, H, W, Cin;
  const float* dy; long ldpix_dy; int Cout;
  float* dw;
  int KH, KW, KK, ph, pw, Ho, Wo, Hp, Wp;
  int R, tiles; long total_rows;
  int x_is_m;          // 1: M side = X channels, N side = dY channels; 0: the other way round
  int Gx, Gy;          // channel groups (of 8) held by the X / dY image (16 on the M side, Np/8 on the N side)
  int Np;              // padded N-side channel count (multiple of 16)
  int nsets;           // accumulator sets (k-steps split over issuers when there are fewer than WISS taps)
  int x3;
  int dbg;
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
// MN-major operands (both): bits 15 / 16 of the instruction descriptor (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint32_t make_idesc_mn(int n) { return make_idesc(n) | (1u << 15) | (1u << 16); }

// 8 channels of one source pixel (zeros outside [0, C))
__device__ __forceinline__ void load_group(const float* src, int c0, int C, bool vec, float (&v)[8]) {
  if (vec && c0 + 8 <= C) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + c0)), b = __ldg(reinterpret_cast<const float4*>(src + c0) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (c0 + i < C) ? __ldg(src + c0 + i) : 0.f;
  }
}

__global__ void __launch_bounds__(WTHREADS, 2) conv_wgrad_shift_kernel(Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // warp index the compiler can prove warp-uniform: the MMA issuers' descriptors then live in uniform registers
  // (otherwise every tcgen05.mma is wrapped in an elect / R2UR.BROADCAST waterfall loop, ~160 cycles per issue)
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t mma_bar = sbase;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 16);
  const int R = p.R, KK = p.KK, Np = p.Np;
  const int x_half = p.Gx * R * 16, y_half = p.Gy * WBM * 16;
  unsigned char* x_hi = smem + WHDR;        // [Gx][R][16 B]
  unsigned char* x_lo = x_hi + x_half;
  unsigned char* y_hi = x_lo + x_half;      // [Gy][128][16 B]
  unsigned char* y_lo = y_hi + y_half;
  const int acc_cols = p.nsets * KK * Np;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)acc_cols) ncols <<= 1;

  if (tid == 0) {
    mbar_init(mma_bar, WISS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 16, ncols);
  // channel groups that no staging pass writes (padding up to 128 M-side / Np N-side channels) are zero for good
  {
    const int total16 = (2 * x_half + 2 * y_half) / 16;
    for (int i = tid; i < total16; i += WTHREADS) reinterpret_cast<uint4*>(x_hi)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int gx = (p.Cin + 7) >> 3, gy = (p.Cout + 7) >> 3;   // staged channel groups
  const bool vec_x = (p.ldpix_x & 3) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
  const bool vec_y = (p.ldpix_dy & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dy) & 15) == 0;
  const int HpWp = p.Hp * p.Wp;

  uint32_t parity = 0;
  int iter = 0;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, parity ^= 1u, ++iter) {
    const long a0 = (long)tile * WBM;
    if (warp < 8) {
      // ---- item = (pixel row, batch of 4 channel groups): one index decomposition and 8 independent 16-byte loads
      //      in flight per item.  X: padded rows [a0, a0 + R); dY: the 128 anchors (anchors in the padding, hp >= Ho or
      //      wp >= Wo, contribute nothing).
      const int bx = (gx + 3) >> 2, by = (gy + 3) >> 2;
      const int items_x = bx * R, items = items_x + by * WBM;
      for (int it = tid; it < items; it += WWORK) {
        const bool is_x = it < items_x;
        const int j = is_x ? it : it - items_x;
        const int rows = is_x ? R : WBM;
        const int r = j % rows, g0 = (j / rows) * 4;
        const int q = (int)a0 + r;
        const float* src = nullptr;
        if (q < (int)p.total_rows) {
          const int n = q / HpWp, rem = q - n * HpWp;
          const int hp = rem / p.Wp, wp = rem - hp * p.Wp;
          if (is_x) {
            const int hs = hp - p.ph, ws = wp - p.pw;
            if (hs >= 0 && hs < p.H && ws >= 0 && ws < p.W) src = p.x + ((long)(n * p.H + hs) * p.W + ws) * p.ldpix_x;
          } else if (hp < p.Ho && wp < p.Wo) {
            src = p.dy + ((long)(n * p.Ho + hp) * p.Wo + wp) * p.ldpix_dy;
          }
        }
        const int C = is_x ? p.Cin : p.Cout, ng = is_x ? gx : gy;
        const bool vec = is_x ? vec_x : vec_y;
        float v[4][8];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (src != nullptr && g0 + b < ng) {
            load_group(src, (g0 + b) * 8, C, vec, v[b]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[b][i] = 0.f;
          }
        }
        unsigned char* img_hi = is_x ? x_hi : y_hi;
        const int half = is_x ? x_half : y_half;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (g0 + b < ng) {
            uint4 hi, lo;
            pack8w(v[b], hi, lo);
            const int off = ((g0 + b) * rows + r) * 16;
            *reinterpret_cast<uint4*>(img_hi + off) = hi;
            if (p.x3) *reinterpret_cast<uint4*>(img_hi + half + off) = lo;
          }
        }
      }
      fence_async_smem();
    }
    __syncthreads();
    if (warp_u >= 8 && elect_one()) {
      const int iss = warp_u - 8;
      tc_fence_after();
      const uint32_t idesc = make_idesc_mn(Np);
      // MN-major, no swizzle: SBO = stride between channel groups, LBO = stride between 8-pixel k-blocks (linear)
      const uint32_t x_sbo = (uint32_t)R * 16, y_sbo = (uint32_t)WBM * 16;
      const uint32_t sx = smem_u32(x_hi), sy = smem_u32(y_hi);
      // work items: nsets == 1 -> this issuer owns taps iss, iss + WISS, ... (all 8 k-steps);
      //             nsets  > 1 -> issuers iss < nsets own k-steps iss, iss + nsets, ... of every tap (set = iss)
      const int t_beg = p.nsets == 1 ? iss : 0, t_inc = p.nsets == 1 ? WISS : 1;
      const int k_beg = p.nsets == 1 ? 0 : iss, k_inc = p.nsets == 1 ? 1 : p.nsets;
      const int set = p.nsets == 1 ? 0 : iss;
      if (p.nsets == 1 || iss < p.nsets) {
        // k-step outermost: consecutive MMAs of one issuer go to different accumulators (an MMA into the accumulator
        // of its predecessor waits for that one to drain)
        const bool first_tile = iter == 0;
        for (int ks = k_beg; ks < WBM / 16; ks += k_inc) {
          const bool first = first_tile && ks == k_beg;
          const uint32_t yh = sy + (uint32_t)(ks * 16 * 16), yl = yh + (uint32_t)y_half;
          const uint64_t dyh = make_desc(yh, 128, y_sbo), dyl = make_desc(yl, 128, y_sbo);
          for (int t = t_beg; t < KK; t += t_inc) {
            const int shift = (t / p.KW) * p.Wp + (t % p.KW);
            const uint32_t d = tmem_base + (uint32_t)((set * KK + t) * Np);
            const uint32_t xh = sx + (uint32_t)((ks * 16 + shift) * 16), xl = xh + (uint32_t)x_half;
            const uint64_t dxh = make_desc(xh, 128, x_sbo);
            const uint64_t ah = p.x_is_m ? dxh : dyh, bh = p.x_is_m ? dyh : dxh;
            if (p.x3) {
              const uint64_t dxl = make_desc(xl, 128, x_sbo);
              const uint64_t al = p.x_is_m ? dxl : dyl, bl = p.x_is_m ? dyl : dxl;
              mma_bf16(d, al, bh, idesc, first ? 0u : 1u);
              mma_bf16(d, ah, bl, idesc, 1u);
              mma_bf16(d, ah, bh, idesc, 1u);
            } else {
              mma_bf16(d, ah, bh, idesc, first ? 0u : 1u);
            }
          }
        }
      }
      mma_commit(mma_bar);
    }
    mbar_wait(mma_bar, parity);   // the images may be overwritten
    tc_fence_after();
  }

  // ---- epilogue: lane = M-side channel, columns = (set, tap, N-side channel); dW[co][ci][tap] += ...
  if (warp < 8 && iter > 0) {
    const int m = (warp & 3) * 32 + lane;
    const int Cm = p.x_is_m ? p.Cin : p.Cout, Cn = p.x_is_m ? p.Cout : p.Cin;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int nchunks = KK * (Np >> 3);
    for (int ch = (warp >> 2); ch < nchunks; ch += 2) {
      const int t = ch / (Np >> 3), n0 = (ch % (Np >> 3)) * 8;
      if (n0 >= Cn) continue;   // warp-uniform
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int s = 0; s < p.nsets; ++s) {
        float v[8];
        tmem_ld8w(t_lane + (uint32_t)((s * KK + t) * Np + n0), v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
      if (m < Cm && !(p.dbg & 64)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = n0 + i;
          if (n < Cn) {
            const int co = p.x_is_m ? n : m, ci = p.x_is_m ? m : n;
            atomicAdd(p.dw + ((long)co * p.Cin + ci) * KK + t, acc[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

}  // namespace wgrads

static inline int wg_round_up(int a, int b) { return (a + b - 1) / b * b; }

// true (and launches: dw += weight gradient) when the shifted-window kernel applies; false -> implicit-GEMM engine
bool conv_wgrad_shift_launch(const float* dy, long ldpix_dy, const float* x, long ldpix_x, int N, int H, int W, int Cin,
                             float* dw, int Cout, int KH, int KW, int ph, int pw, int Ho, int Wo, void* stream) {
  using namespace wgrads;
  if (umma::g_dbg_flags & 4) return false;  // bring-up switch: force the implicit-GEMM path
  if (ph < 0 || pw < 0) return false;
  const int Hp = H + 2 * ph, Wp = W + 2 * pw;
  if (Hp < KH || Wp < KW || Ho != Hp - KH + 1 || Wo != Wp - KW + 1) return false;
  const int KK = KH * KW;
  // Where it pays (measured on the layer shapes of the G / D nets at 256 clips, tools/bench_wgrad.py): shapes the
  // implicit GEMM would run on the SIMT engine (a side narrower than a tensor-core tile), and multi-tap filters whose
  // Cout*Cin*KK gradient is small enough that the per-CTA red.global epilogue (every CTA adds its whole partial
  // gradient to the same addresses) stays cheap.  1x1 convolutions with wide channels are plain GEMMs already.
  {
    const bool simt_fallback = Cout < 16 || Cin * KK < 16;
    const bool small_multitap = KK >= 3 && Cin >= 16 && (long)Cout * Cin * KK <= 16384;
    if (!(umma::g_dbg_flags & 128) && !simt_fallback && !small_multitap) return false;
  }
  // role assignment: M side <= 128 channels, KK * pad16(N side) <= 512 TMEM columns; prefer the wider side on M
  int x_is_m = -1;
  {
    const bool ok_x = Cin <= 128 && KK * wg_round_up(Cout, 16) <= 512 && wg_round_up(Cout, 16) <= 256;
    const bool ok_y = Cout <= 128 && KK * wg_round_up(Cin, 16) <= 512 && wg_round_up(Cin, 16) <= 256;
    if (ok_x && ok_y) x_is_m = Cin >= Cout ? 1 : 0;
    else if (ok_x) x_is_m = 1;
    else if (ok_y) x_is_m = 0;
    else return false;
  }
  const int Np = wg_round_up(x_is_m ? Cout : Cin, 16);
  const int R = wg_round_up(WBM + (KH - 1) * Wp + (KW - 1), 8);
  const int Gx = x_is_m ? 16 : Np / 8, Gy = x_is_m ? Np / 8 : 16;
  const size_t smem = WHDR + 2 * (size_t)Gx * R * 16 + 2 * (size_t)Gy * WBM * 16;
  if (smem > 220 * 1024 || R > 8192) return false;
  const long total = (long)N * Hp * Wp;
  if (total <= 0 || total > (1L << 30)) return false;
  Params p;
  p.x = x; p.ldpix_x = ldpix_x; p.N = N; p.H = H; p.W = W; p.Cin = Cin;
  p.dy = dy; p.ldpix_dy = ldpix_dy; p.Cout = Cout; p.dw = dw;
  p.KH = KH; p.KW = KW; p.KK = KK; p.ph = ph; p.pw = pw; p.Ho = Ho; p.Wo = Wo; p.Hp = Hp; p.Wp = Wp;
  p.R = R; p.total_rows = total; p.tiles = (int)((total + WBM - 1) / WBM);
  p.x_is_m = x_is_m; p.Gx = Gx; p.Gy = Gy; p.Np = Np;
  int nsets = 1;
  if (KK < WISS) {
    nsets = WISS;
    while (nsets > 1 && nsets * KK * Np > 512) nsets >>= 1;
  }
  p.nsets = nsets;
  p.x3 = umma::g_precision == 0 ? 1 : 0;
  p.dbg = umma::g_dbg_flags;
  auto kfn = &conv_wgrad_shift_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  // two CTAs per SM (one stages while the other's MMAs run) when shared memory and TMEM columns allow; otherwise the
  // dynamic shared memory request is padded so that a second CTA (whose tcgen05.alloc would block) never co-resides.
  // Every CTA ends with Cout*Cin*KK atomics, so it should own several tiles.
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(nsets * KK * Np)) ncols <<= 1;
  const int per_sm = (smem + 1024 <= (size_t)110 * 1024 && ncols <= 256) ? 2 : 1;
  size_t smem_req = smem;
  if (per_sm == 1 && smem_req < (size_t)116 * 1024) smem_req = (size_t)116 * 1024;
  if (per_sm == 2 && smem_req < (size_t)78 * 1024) smem_req = (size_t)78 * 1024;   // never three
  int grid = p.tiles;
  if (grid > s2ag_sm_count() * per_sm) grid = s2ag_sm_count() * per_sm;
  if (grid < 1) grid = 1;
  S2AG_LAUNCH(kfn, grid, WTHREADS, smem_req, stream, p);
  return true;
}

}  // namespace s2ag
