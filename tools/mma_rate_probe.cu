// Development aid: issue/execution rate of tcgen05.mma (kind::f16, bf16 operands, M = 128, K = 16) on one SM as a
// function of N, of the B operand's major-ness and of how many TMEM accumulators the instruction stream rotates over.
// One CTA, one elected thread issues R MMAs back to back and commits; clock64 around issue..commit-completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include tools/mma_rate_probe.cu -o tools/_build/mma_rate_probe
#include "../speech2affective_gestures_b200/csrc/gemm_umma.cuh"
#include <cstdarg>
unsigned long long g_s2ag_launches = 0;
void s2ag_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
namespace s2ag { int g_engine = 0; namespace umma { int g_precision = 0; int g_dbg_flags = 0; } }
using namespace s2ag::umma;

template <int N, int NACC, bool BMN, int AROWS = 128, int ASHIFT = 0>
__global__ void __launch_bounds__(128, 1) probe(int R, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + 16);
  unsigned char* a = smem + 1024;              // A: K-major [k-chunk][128][16], 16 k-chunks
  unsigned char* b = a + 16 * 144 * 16;        // B: K-major [k-chunk][256][16] or MN-major [group of 8][k rows][16]
  for (int i = tid; i < (16 * 144 * 16 + 16 * 256 * 16) / 16; i += 128) reinterpret_cast<uint4*>(a)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(sbase + 16, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  if (warp_u == 1 && elect_one()) {
    const uint32_t idesc = make_idesc(N) | (BMN ? (1u << 16) : 0u);
    const uint64_t da0 = make_desc(smem_u32(a) + ASHIFT * 16, AROWS * 16, 128);   // AROWS: chunk stride in rows, ASHIFT: row shift
    const uint64_t db0 = BMN ? make_desc(smem_u32(b), 128, 128 * 16) : make_desc(smem_u32(b), 256 * 16, 128);
    constexpr uint32_t astep = (2 * AROWS * 16) >> 4, bstep = BMN ? (256 >> 4) : ((2 * 256 * 16) >> 4);
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int it = 0; it < R / 8; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_bf16(tb + (uint32_t)((ks % NACC) * N), da0 + (uint64_t)(ks * astep), db0 + (uint64_t)(ks * bstep), idesc,
                   (it > 0 || ks >= NACC) ? 1u : 0u);
      }
      const long long t1 = clock64();
      mma_commit(bar);
      mbar_wait(bar, (uint32_t)(rep & 1));
      const long long t2 = clock64();
      out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int NACC, bool BMN, int AROWS = 128, int ASHIFT = 0>
static void run(long long* d, int R) {
  auto k = &probe<N, NACC, BMN, AROWS, ASHIFT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  k<<<1, 128, 1024 + 16 * 144 * 16 + 16 * 256 * 16>>>(R, d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
  long long h[6];
  cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("  B %s  N=%3d  nacc=%d  A chunk stride %d rows, start row %d : issue %6.1f  total %6.1f\n", BMN ? "MN-major" : "K-major ", N, NACC, AROWS, ASHIFT, h[4] / (double)R, h[5] / (double)R);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int R = 64;
  printf("tcgen05.mma kind::f16 M=128 K=16, %d MMAs by one thread: cycles per MMA (issue loop | until commit completes)\n", R);
  run<16, 1, false>(d, R); run<32, 1, false>(d, R); run<32, 2, false>(d, R); run<32, 4, false>(d, R);
  run<48, 1, false>(d, R); run<48, 4, false>(d, R); run<64, 1, false>(d, R); run<64, 2, false>(d, R); run<64, 4, false>(d, R);
  run<128, 1, false>(d, R); run<128, 2, false>(d, R); run<256, 1, false>(d, R); run<256, 2, false>(d, R);
  run<144, 1, false>(d, R); run<160, 1, false>(d, R); run<160, 1, false, 137, 0>(d, R); run<160, 1, false, 137, 2>(d, R); run<160, 1, false, 136, 2>(d, R); run<160, 1, false, 136, 8>(d, R); run<64, 1, false, 132, 1>(d, R);
  run<32, 1, true>(d, R); run<32, 4, true>(d, R); run<64, 1, true>(d, R); run<64, 2, true>(d, R); run<128, 2, true>(d, R); run<256, 2, true>(d, R);
  return 0;
}
