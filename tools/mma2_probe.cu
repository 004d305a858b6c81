// Development aid: semantics check of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, kind::f16, bf16 operands,
// K-major no-swizzle layouts as used by this library).  Each CTA of a 2-CTA cluster holds 128 rows of A and N/2 rows of
// B at the same shared-memory offsets; the leader (cluster rank 0) issues ONE MMA; both CTAs read their 128 x N
// accumulator from their own TMEM and compare with the host result for the hypothesis "CTA c supplies B rows
// [c N/2, (c+1) N/2)".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include tools/mma2_probe.cu -o tools/_build/mma2_probe
#include "../speech2affective_gestures_b200/csrc/gemm_umma.cuh"
#include <cstdarg>
#include <vector>
unsigned long long g_s2ag_launches = 0;
void s2ag_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
namespace s2ag { int g_engine = 0; namespace umma { int g_precision = 0; int g_dbg_flags = 0; } }
using namespace s2ag::umma;

constexpr int N = 32, K = 16;
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(const float* a_all, const float* b_all, float* d_all) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ctarank();
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + 16);
  __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(smem + 1024);              // [k-chunk 2][128][8]
  __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(smem + 1024 + 2 * 128 * 16);  // [k-chunk 2][N/2][8]
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    a[((k / 8) * 128 + r) * 8 + (k % 8)] = __float2bfloat16(a_all[(rank * 128 + r) * K + k]);
  }
  for (int i = tid; i < (N / 2) * K; i += 128) {
    const int n = i / K, k = i % K;
    b[((k / 8) * (N / 2) + n) * 8 + (k % 8)] = __float2bfloat16(b_all[(rank * (N / 2) + n) * K + k]);
  }
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + 16), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' operands are staged and their barriers initialised
  tc_fence_after();
  const uint32_t tb = *slot;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  if (rank == 0 && warp_u == 1 && elect_one()) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(a), 128 * 16, 128), db = make_desc(smem_u32(b), (N / 2) * 16, 128);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  {
    uint32_t r[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16), r);
    for (int n = 0; n < N; ++n) d_all[((rank * 128) + warp * 32 + lane) * N + n] = __uint_as_float(r[n]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32u) : "memory");
}


// rate: R back-to-back pair MMAs (M = 256, N = NN, K = 16) issued by the leader; cycles per MMA until the commit completes
template <int NN, int CEVERY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate(int R, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = ctarank();
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + 16);
  unsigned char* a = smem + 1024;                    // 16 k-chunks x 128 rows x 16 B
  unsigned char* b = a + 16 * 128 * 16;              // 16 k-chunks x NN/2 rows x 16 B
  for (int i = tid; i < (16 * 128 * 16 + 16 * 128 * 16) / 16; i += 128) reinterpret_cast<uint4*>(a)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 32, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + 16), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = *slot;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  if (rank == 0 && warp_u == 1 && elect_one()) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t da0 = make_desc(smem_u32(a), 128 * 16, 128), db0 = make_desc(smem_u32(b), (NN / 2) * 16, 128);
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int it = 0; it < R; ++it) {
        const int ks = it & 7;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tb), "l"(da0 + (uint64_t)(ks * ((2 * 128 * 16) >> 4))), "l"(db0 + (uint64_t)(ks * ((2 * (NN / 2) * 16) >> 4))),
                       "r"(idesc), "r"(it ? 1u : 0u) : "memory");
        if (CEVERY > 0 && (it % CEVERY) == CEVERY - 1)
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                       ::"r"(bar + 32), "h"((uint16_t)3) : "memory");
      }
      const long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(bar), "h"((uint16_t)3) : "memory");
      mbar_wait(bar, (uint32_t)(rep & 1));
      const long long t2 = clock64();
      out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0;
    }
  } else if (rank == 1 && tid == 0) {
    for (int rep = 0; rep < 3; ++rep) mbar_wait(bar, (uint32_t)(rep & 1));
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}
template <int NN, int CEVERY>
static void run_rate(long long* d) {
  const int R = 64;
  cudaFuncSetAttribute(rate<NN, CEVERY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  rate<NN, CEVERY><<<2, 128, 1024 + 2 * 16 * 128 * 16>>>(R, d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("rate kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
  long long h[6];
  cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("  cta_group::2 M=256 N=%3d K=16, multicast commit every %d MMAs: issue %6.1f  total %6.1f cycles per MMA\n", NN, CEVERY, h[4] / (double)R, h[5] / (double)R);
}

int main() {
  std::vector<float> a(256 * K), b(N * K), d(256 * N, -1.f);
  for (size_t i = 0; i < a.size(); ++i) a[i] = (float)((int)(i * 7 % 13) - 6);
  for (size_t i = 0; i < b.size(); ++i) b[i] = (float)((int)(i * 5 % 11) - 5);
  float *da, *db, *dd;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xff, d.size() * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<<<2, 128, 32 * 1024>>>(da, db, dd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < 256; ++r)
    for (int n = 0; n < N; ++n) {
      float ref = 0.f;
      for (int k = 0; k < K; ++k) ref += a[r * K + k] * b[n * K + k];
      if (d[r * N + n] != ref) { if (bad < 8) printf("mismatch row %d col %d: got %g want %g\n", r, n, d[r * N + n], ref); ++bad; }
    }
  printf("tcgen05.mma.cta_group::2 M=256 N=%d K=%d: %s (%d mismatches of %d)\n", N, K, bad ? "MISMATCH" : "OK: CTA c holds A rows [128c,128c+128) and B rows [cN/2,(c+1)N/2); each CTA's TMEM gets its 128 x N block", bad, 256 * N);
  long long* dt; cudaMalloc(&dt, 64);
  run_rate<32, 0>(dt); run_rate<64, 0>(dt); run_rate<128, 0>(dt); run_rate<144, 0>(dt); run_rate<160, 0>(dt); run_rate<256, 0>(dt);
  run_rate<160, 3>(dt); run_rate<160, 6>(dt); run_rate<160, 12>(dt);
  return bad != 0;
}
