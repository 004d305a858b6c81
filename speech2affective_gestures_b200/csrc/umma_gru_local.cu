// CTA-LOCAL bi-GRU recurrence / BPTT for small hidden sizes (H <= 64: the AffDiscriminator's nn.GRU(8, 64, 4 layers),
// net/multimodal_context_net_v2.py:558-560) on tcgen05 (sm_100a).
//
// The slice-parallel persistent kernels of umma_gru.cu split the hidden units over CTAs and exchange h (or the BPTT
// partial products) through L2 every time step: flags, gpu-scope fences, TMA fetches -- a ~3.4 us (forward) / ~6 us
// (BPTT) latency chain per step that does not shrink with H.  For H <= 64 the whole W_hh (3H x H) fits one CTA: here a
// CTA owns ALL hidden units of its 128 clips and direction, so a time step is
//     forward :  [128 x Kp] h image (shared memory)  x  W_hh^T  ->  [128 x 3Kp] gate pre-activations in TMEM
//     BPTT    :  [128 x 3Kp] dgh image (shared memory) x  W_hh   ->  [128 x Kp] carry in TMEM
// followed by the gate math of the 8 worker warps, which write the next operand image straight back to shared memory:
// no inter-CTA traffic, no fences, two mbarriers per step.  grid = (batch tiles of 128, 2 directions).
// Same data contracts as umma_gru.cu (gi / gates / dgi / dgh layouts), so the callers in gru.cu switch freely.
#include "s2ag.h"
#include "gemm_umma.cuh"

namespace s2ag {
namespace grul {

using namespace s2ag::umma;

constexpr int LBM = 128, LTHREADS = 288, LHDR = 128, LMAXH = 64, LMAXCH = 4;  // 8 worker warps + 1 MMA-issuer warp

struct FwdP {
  const float* gi;                       // [B*T][2][3H] input projections incl. b_ih
  const float* whh; long whh_dstride;    // [3H][H] per direction
  const float* bhh; long bhh_dstride;
  float* out;                            // [B][T][2H]
  float* gates;                          // [T][2][4][H][B] (r, z, n, W_hn h + b_hn) or NULL
  int B, T, H, Kp, x3;
};
struct BwdP {
  const float* dout; long lddout; int dir_stride;
  const float* out; const float* gates;
  const float* whh; long whh_dstride;
  float* dgi; float* dgh;                // [B*T][2][3H]
  int B, T, H, Kp, x3;
};

__device__ __forceinline__ void pack8l(const float (&v)[8], uint4& hi, uint4& lo) {
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
__device__ __forceinline__ void tmem_ld8_nw(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 8 consecutive floats (zeros beyond `valid`); 16-byte loads when `vec`
__device__ __forceinline__ void ld8(const float* q, int valid, bool vec, float (&v)[8]) {
  if (vec && valid >= 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(q)), b = __ldg(reinterpret_cast<const float4*>(q) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < valid ? __ldg(q + i) : 0.f;
  }
}
__device__ __forceinline__ void st8(float* q, int valid, bool vec, const float (&v)[8]) {
  if (vec && valid >= 8) {
    reinterpret_cast<float4*>(q)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(q)[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < valid) q[i] = v[i];
  }
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(LTHREADS, 1) gru_local_fwd_kernel(FwdP p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int bt = blockIdx.x, dir = blockIdx.y;
  const int H = p.H, T = p.T, B = p.B, Kp = p.Kp, Kc = Kp >> 3, NP = 3 * Kp;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t mma_bar = sbase, a_bar = sbase + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 16);
  const int w_half = Kc * NP * 16, a_half = Kc * LBM * 16;
  unsigned char* w_hi = smem + LHDR;            // B operand: [k-chunk][n = g*Kp + j][16 B]
  unsigned char* w_lo = w_hi + w_half;
  unsigned char* a_hi = w_lo + w_half;          // A operand: h image [k-chunk][clip row][16 B]
  unsigned char* a_lo = a_hi + a_half;
  float* bh_s = reinterpret_cast<float*>(a_lo + a_half);   // b_hh, [3][Kp]
  uint32_t ncols = 32;
  while (ncols < (uint32_t)NP) ncols <<= 1;

  if (tid == 0) {
    mbar_init(mma_bar, 1);
    mbar_init(a_bar, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 16, ncols);
  const float* whh = p.whh + dir * p.whh_dstride;
  for (int idx = tid; idx < Kc * NP; idx += LTHREADS) {
    const int n = idx % NP, kc = idx / NP;
    const int g = n / Kp, j = n % Kp;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kc * 8 + i;
      v[i] = (j < H && k < H) ? __ldg(whh + ((long)g * H + j) * H + k) : 0.f;
    }
    uint4 hi, lo;
    pack8l(v, hi, lo);
    *reinterpret_cast<uint4*>(w_hi + idx * 16) = hi;
    *reinterpret_cast<uint4*>(w_lo + idx * 16) = lo;
  }
  for (int idx = tid; idx < NP; idx += LTHREADS) {
    const int g = idx / Kp, j = idx % Kp;
    bh_s[idx] = j < H ? __ldg(p.bhh + dir * p.bhh_dstride + g * H + j) : 0.f;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_u == 8) {
    // ================================ MMA issuer: gate pre-activations of step s from the h_{s-1} image
    if (elect_one()) {
      const uint32_t idesc = make_idesc(NP);
      const uint32_t a_lbo = LBM * 16, w_lbo = (uint32_t)NP * 16;
      const uint32_t sa = smem_u32(a_hi), sw = smem_u32(w_hi);
      for (int s = 1; s < T; ++s) {
        mbar_wait(a_bar, (uint32_t)((s - 1) & 1));
        tc_fence_after();
        for (int kk = 0; kk < (Kp >> 4); ++kk) {
          const uint32_t ah = sa + kk * 2 * a_lbo, al = ah + (uint32_t)a_half;
          const uint32_t wh = sw + kk * 2 * w_lbo, wl = wh + (uint32_t)w_half;
          const uint64_t dah = make_desc(ah, a_lbo, 128), dwh = make_desc(wh, w_lbo, 128);
          if (p.x3) {
            mma_bf16(tmem_base, make_desc(al, a_lbo, 128), dwh, idesc, kk > 0 ? 1u : 0u);
            mma_bf16(tmem_base, dah, make_desc(wl, w_lbo, 128), idesc, 1u);
            mma_bf16(tmem_base, dah, dwh, idesc, 1u);
          } else {
            mma_bf16(tmem_base, dah, dwh, idesc, kk > 0 ? 1u : 0u);
          }
        }
        mma_commit(mma_bar);
      }
    }
  } else {
    // ================================ workers: thread = clip row x chunks (wg, wg + 2, ...) of 8 hidden units
    const int row = (warp & 3) * 32 + lane;
    const int b = bt * LBM + row;
    const int wg = warp >> 2;
    const bool b_ok = b < B;
    const bool vec = (H & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.gi) | reinterpret_cast<uintptr_t>(p.out)) & 15) == 0;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float h_own[LMAXCH][8];
#pragma unroll
    for (int i = 0; i < LMAXCH; ++i)
#pragma unroll
      for (int u = 0; u < 8; ++u) h_own[i][u] = 0.f;

    for (int s = 0; s < T; ++s) {
      const int t = dir == 0 ? s : T - 1 - s;
      // gi of this step for all chunks: issued before the accumulator wait (independent of the recurrence)
      float gr[LMAXCH][8], gz[LMAXCH][8], gn[LMAXCH][8];
      const float* g0 = p.gi + ((long)(b_ok ? b : 0) * T + t) * 6 * H + (long)dir * 3 * H;
#pragma unroll
      for (int i = 0; i < LMAXCH; ++i) {
        const int c = wg + 2 * i;
        const int valid = (c < Kc && b_ok) ? (H - c * 8 < 8 ? H - c * 8 : 8) : 0;
        ld8(g0 + c * 8, valid, vec, gr[i]);
        ld8(g0 + H + c * 8, valid, vec, gz[i]);
        ld8(g0 + 2 * H + c * 8, valid, vec, gn[i]);
      }
      if (s > 0) {
        mbar_wait(mma_bar, (uint32_t)((s - 1) & 1));
        tc_fence_after();
      }
#pragma unroll
      for (int i = 0; i < LMAXCH; ++i) {
        const int c = wg + 2 * i;
        if (c < Kc) {   // warp-uniform
          float ar[8], az[8], an[8];
          if (s > 0) {
            tmem_ld8_nw(t_lane + (uint32_t)(0 * Kp + c * 8), ar);
            tmem_ld8_nw(t_lane + (uint32_t)(1 * Kp + c * 8), az);
            tmem_ld8_nw(t_lane + (uint32_t)(2 * Kp + c * 8), an);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) ar[u] = az[u] = an[u] = 0.f;
          }
          float rr[8], zz[8], nn[8], gh[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = c * 8 + u;
            const bool ok = b_ok && j < H;
            gh[u] = an[u] + bh_s[2 * Kp + j];
            rr[u] = __fdividef(1.f, 1.f + __expf(-(gr[i][u] + ar[u] + bh_s[j])));
            zz[u] = __fdividef(1.f, 1.f + __expf(-(gz[i][u] + az[u] + bh_s[Kp + j])));
            nn[u] = 1.f - __fdividef(2.f, __expf(2.f * (gn[i][u] + rr[u] * gh[u])) + 1.f);
            const float h = (1.f - zz[u]) * nn[u] + zz[u] * h_own[i][u];
            h_own[i][u] = ok ? h : 0.f;
          }
          if (s + 1 < T) {  // operand image of h_s (the MMAs that read h_{s-1} have completed)
            uint4 hi, lo;
            pack8l(h_own[i], hi, lo);
            *reinterpret_cast<uint4*>(a_hi + (c * LBM + row) * 16) = hi;
            if (p.x3) *reinterpret_cast<uint4*>(a_lo + (c * LBM + row) * 16) = lo;
          }
          if (b_ok) {
            const int valid = H - c * 8 < 8 ? H - c * 8 : 8;
            st8(p.out + ((long)b * T + t) * 2 * H + (long)dir * H + c * 8, valid, vec, h_own[i]);
            if (p.gates) {
              float* gs = p.gates + ((((long)t * 2 + dir) * 4) * H + c * 8) * B + b;
              const long gstride = (long)H * B;
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                if (u < valid) {
                  gs[(long)u * B] = rr[u]; gs[gstride + (long)u * B] = zz[u]; gs[2 * gstride + (long)u * B] = nn[u];
                  gs[3 * gstride + (long)u * B] = gh[u];
                }
              }
            }
          }
        }
      }
      if (s + 1 < T) {
        tc_fence_before();      // accumulator reads done before the next step's MMAs overwrite it
        fence_async_smem();     // image stores -> async proxy (tensor core)
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(a_bar);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------ BPTT
//   dh = dout + carry;  dn = dh(1-z)(1-n^2);  dz = dh(h_prev - n)z(1-z);  dr = dn*ghn*r(1-r)
//   dgi = (dr,dz,dn), dgh = (dr,dz,dn*r);  carry' = dh*z + dgh @ W_hh
__global__ void __launch_bounds__(LTHREADS, 1) gru_local_bwd_kernel(BwdP p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int bt = blockIdx.x, dir = blockIdx.y;
  const int H = p.H, T = p.T, B = p.B, Kp = p.Kp, Kc = Kp >> 3, NP = 3 * Kp, NCc = NP >> 3;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t mma_bar = sbase, a_bar = sbase + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 16);
  const int w_half = NCc * Kp * 16, a_half = NCc * LBM * 16;
  unsigned char* w_hi = smem + LHDR;            // B operand: [k-chunk over c = g*Kp + j][n = k_out][16 B]
  unsigned char* w_lo = w_hi + w_half;
  unsigned char* a_hi = w_lo + w_half;          // A operand: dgh image [k-chunk][clip row][16 B]
  unsigned char* a_lo = a_hi + a_half;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)Kp) ncols <<= 1;

  if (tid == 0) {
    mbar_init(mma_bar, 1);
    mbar_init(a_bar, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + 16, ncols);
  const float* whh = p.whh + dir * p.whh_dstride;
  for (int idx = tid; idx < NCc * Kp; idx += LTHREADS) {
    const int n = idx % Kp, kc = idx / Kp;      // kc: chunk of 8 gate rows c
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = kc * 8 + i, g = c / Kp, j = c % Kp;
      v[i] = (j < H && n < H) ? __ldg(whh + ((long)g * H + j) * H + n) : 0.f;
    }
    uint4 hi, lo;
    pack8l(v, hi, lo);
    *reinterpret_cast<uint4*>(w_hi + idx * 16) = hi;
    *reinterpret_cast<uint4*>(w_lo + idx * 16) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_u == 8) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(Kp);
      const uint32_t a_lbo = LBM * 16, w_lbo = (uint32_t)Kp * 16;
      const uint32_t sa = smem_u32(a_hi), sw = smem_u32(w_hi);
      for (int step = 0; step + 1 < T; ++step) {   // the last step (forward index 0) has no predecessor to feed
        mbar_wait(a_bar, (uint32_t)(step & 1));
        tc_fence_after();
        for (int kk = 0; kk < (NP >> 4); ++kk) {
          const uint32_t ah = sa + kk * 2 * a_lbo, al = ah + (uint32_t)a_half;
          const uint32_t wh = sw + kk * 2 * w_lbo, wl = wh + (uint32_t)w_half;
          const uint64_t dah = make_desc(ah, a_lbo, 128), dwh = make_desc(wh, w_lbo, 128);
          if (p.x3) {
            mma_bf16(tmem_base, make_desc(al, a_lbo, 128), dwh, idesc, kk > 0 ? 1u : 0u);
            mma_bf16(tmem_base, dah, make_desc(wl, w_lbo, 128), idesc, 1u);
            mma_bf16(tmem_base, dah, dwh, idesc, 1u);
          } else {
            mma_bf16(tmem_base, dah, dwh, idesc, kk > 0 ? 1u : 0u);
          }
        }
        mma_commit(mma_bar);
      }
    }
  } else {
    const int row = (warp & 3) * 32 + lane;
    const int b = bt * LBM + row;
    const int wg = warp >> 2;
    const bool b_ok = b < B;
    const bool vec = (H & 3) == 0 && (p.lddout & 3) == 0 && (p.dir_stride & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(p.dout) | reinterpret_cast<uintptr_t>(p.out) |
                       reinterpret_cast<uintptr_t>(p.dgi) | reinterpret_cast<uintptr_t>(p.dgh)) & 15) == 0;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float carry[LMAXCH][8];
#pragma unroll
    for (int i = 0; i < LMAXCH; ++i)
#pragma unroll
      for (int u = 0; u < 8; ++u) carry[i][u] = 0.f;

    for (int step = 0; step < T; ++step) {
      const int fs = T - 1 - step;                       // forward step index being differentiated
      const int t = dir == 0 ? fs : T - 1 - fs;
      const int tprev = dir == 0 ? t - 1 : t + 1;
      const long rowi = (long)(b_ok ? b : 0) * T + t;
#pragma unroll
      for (int i = 0; i < LMAXCH; ++i) {
        const int c = wg + 2 * i;
        if (c < Kc) {   // warp-uniform
          const int valid = b_ok ? (H - c * 8 < 8 ? H - c * 8 : 8) : 0;
          float dh[8], hp[8];
          ld8(p.dout + rowi * p.lddout + (long)dir * p.dir_stride + c * 8, valid, vec, dh);
          if (fs > 0) {
            ld8(p.out + ((long)(b_ok ? b : 0) * T + tprev) * 2 * H + (long)dir * H + c * 8, valid, vec, hp);
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) hp[u] = 0.f;
          }
          const float* gs = p.gates + ((((long)t * 2 + dir) * 4) * H + c * 8) * B + (b_ok ? b : 0);
          const long gstride = (long)H * B;
          float dr[8], dz[8], dn[8], dnr[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const bool ok = u < valid;
            const float r = ok ? __ldg(gs + (long)u * B) : 0.f;
            const float z = ok ? __ldg(gs + gstride + (long)u * B) : 0.f;
            const float n = ok ? __ldg(gs + 2 * gstride + (long)u * B) : 0.f;
            const float ghn = ok ? __ldg(gs + 3 * gstride + (long)u * B) : 0.f;
            const float d = dh[u] + carry[i][u];
            dn[u] = d * (1.f - z) * (1.f - n * n);
            dz[u] = d * (hp[u] - n) * z * (1.f - z);
            dr[u] = dn[u] * ghn * r * (1.f - r);
            dnr[u] = dn[u] * r;
            carry[i][u] = d * z;     // + (dgh @ W_hh) below
          }
          if (b_ok) {
            float* a = p.dgi + (rowi * 2 + dir) * 3 * H + c * 8;
            float* cg = p.dgh + (rowi * 2 + dir) * 3 * H + c * 8;
            st8(a, valid, vec, dr); st8(a + H, valid, vec, dz); st8(a + 2 * H, valid, vec, dn);
            st8(cg, valid, vec, dr); st8(cg + H, valid, vec, dz); st8(cg + 2 * H, valid, vec, dnr);
          }
          if (fs > 0) {   // A operand chunk of gate g: g*Kc + c
            uint4 hi, lo;
            pack8l(dr, hi, lo);
            *reinterpret_cast<uint4*>(a_hi + ((0 * Kc + c) * LBM + row) * 16) = hi;
            if (p.x3) *reinterpret_cast<uint4*>(a_lo + ((0 * Kc + c) * LBM + row) * 16) = lo;
            pack8l(dz, hi, lo);
            *reinterpret_cast<uint4*>(a_hi + ((1 * Kc + c) * LBM + row) * 16) = hi;
            if (p.x3) *reinterpret_cast<uint4*>(a_lo + ((1 * Kc + c) * LBM + row) * 16) = lo;
            pack8l(dnr, hi, lo);
            *reinterpret_cast<uint4*>(a_hi + ((2 * Kc + c) * LBM + row) * 16) = hi;
            if (p.x3) *reinterpret_cast<uint4*>(a_lo + ((2 * Kc + c) * LBM + row) * 16) = lo;
          }
        }
      }
      if (fs > 0) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(a_bar);
        mbar_wait(mma_bar, (uint32_t)(step & 1));
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < LMAXCH; ++i) {
          const int c = wg + 2 * i;
          if (c < Kc) {
            float v[8];
            tmem_ld8_nw(t_lane + (uint32_t)(c * 8), v);
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 8; ++u) carry[i][u] += v[u];
          }
        }
        tc_fence_before();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

static inline int kp_of(int H) { return (H + 15) / 16 * 16; }
static inline size_t fwd_smem(int H) {
  const int Kp = kp_of(H), Kc = Kp / 8, NP = 3 * Kp;
  return LHDR + 2 * (size_t)Kc * NP * 16 + 2 * (size_t)Kc * LBM * 16 + (size_t)NP * 4 + 16;
}
static inline size_t bwd_smem(int H) {
  const int Kp = kp_of(H), NCc = 3 * Kp / 8;
  return LHDR + 2 * (size_t)NCc * Kp * 16 + 2 * (size_t)NCc * LBM * 16;
}

}  // namespace grul

// MEASURED (B200, AffDiscriminator layers, 256-512 clips): correct (the whole GPU suite passes on this path) but SLOWER
// than the slice-parallel kernels -- 20.8 vs 15.6 ms/step: with all 64 hidden units of 128 clips on ONE SM the gate
// math (8192 cells x 6 MUFU ops per step) and 12 N=192 MMAs cost more than the exchange they save, and only 4-8 SMs
// work.  Kept as an opt-in experiment (s2ag_debug_flags bit 9); the productive variant is a 4-CTA cluster exchanging
// h through distributed shared memory (next round).
bool gru_local_supported(int H) { return H >= 8 && H <= grul::LMAXH && (umma::g_dbg_flags & 512) != 0; }

int gru_local_fwd(const float* gi, const float* whh_f, long whh_dstride, const float* bhh_f, long bhh_dstride, float* out,
                  float* gates, int B, int T, int H, int x3, void* stream) {
  using namespace grul;
  if (!gru_local_supported(H)) return S2AG_ERR_UNSUPPORTED;
  auto kfn = &gru_local_fwd_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem(LMAXH));
    attr_set = true;
  }
  FwdP p;
  p.gi = gi; p.whh = whh_f; p.whh_dstride = whh_dstride; p.bhh = bhh_f; p.bhh_dstride = bhh_dstride;
  p.out = out; p.gates = gates; p.B = B; p.T = T; p.H = H; p.Kp = kp_of(H); p.x3 = x3;
  S2AG_LAUNCH(kfn, dim3((B + LBM - 1) / LBM, 2), LTHREADS, fwd_smem(H), stream, p);
  return S2AG_OK;
}

int gru_local_bwd(const float* dout, long lddout, int dir_stride, const float* out, const float* gates,
                  const float* whh_f, long whh_dstride, float* dgi, float* dgh, int B, int T, int H, int x3,
                  void* stream) {
  using namespace grul;
  if (!gru_local_supported(H)) return S2AG_ERR_UNSUPPORTED;
  auto kfn = &gru_local_bwd_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem(LMAXH));
    attr_set = true;
  }
  BwdP p;
  p.dout = dout; p.lddout = lddout; p.dir_stride = dir_stride; p.out = out; p.gates = gates;
  p.whh = whh_f; p.whh_dstride = whh_dstride; p.dgi = dgi; p.dgh = dgh;
  p.B = B; p.T = T; p.H = H; p.Kp = kp_of(H); p.x3 = x3;
  S2AG_LAUNCH(kfn, dim3((B + LBM - 1) / LBM, 2), LTHREADS, bwd_smem(H), stream, p);
  return S2AG_OK;
}

}  // namespace s2ag
