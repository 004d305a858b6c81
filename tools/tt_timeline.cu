// Development aid: clock64 marks of CTA (0,0,0) of the two-TMA contraction kernel on the GRU projection shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -DS2AG_TT_TIMELINE tools/tt_timeline.cu -o tools/_build/tt_timeline
#include <cstdarg>
#include <cstdlib>
#include <vector>
#include "../speech2affective_gestures_b200/csrc/gemm.cuh"
unsigned long long g_s2ag_launches = 0;
void s2ag_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
namespace s2ag { int g_engine = 0; namespace umma { int g_precision = 0; int g_dbg_flags = 0; }
static void* g_buf = nullptr; static long g_bytes = 0;
void* scratch_get(void*, long bytes) { return bytes <= g_bytes ? g_buf : nullptr; } }
using namespace s2ag;

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 8704, N = argc > 2 ? atoi(argv[2]) : 1800, K = argc > 3 ? atoi(argv[3]) : 600;
  umma::g_precision = argc > 4 ? atoi(argv[4]) : 0;
  float *x, *w, *y, *bias;
  cudaMalloc(&x, (size_t)M * K * 4); cudaMalloc(&w, (size_t)N * K * 4); cudaMalloc(&y, (size_t)M * N * 4); cudaMalloc(&bias, N * 4);
  cudaMemset(x, 0, (size_t)M * K * 4); cudaMemset(w, 0, (size_t)N * K * 4); cudaMemset(bias, 0, N * 4);
  g_bytes = 192L << 20; cudaMalloc(&g_buf, g_bytes);
  LdPlain<true> a{x, (long)K, 1, 0}, b{w, (long)K, 1, 0};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    launch_gemm(a, b, make_epi(y, (long)N, bias, 0, 0.f, 0), M, N, K, 1, 1, nullptr);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long tl[4][32];
  cudaMemcpyFromSymbol(tl, umma::g_tt_tl, sizeof(tl));
  const long long t0 = tl[3][0];
  printf("M=%d N=%d K=%d precision %d: %.1f us (packs + kernel)\n", M, N, K, umma::g_precision, ms * 1e3);
  printf("main loop done %lld, kernel end %lld (cycles after the prologue)\n", tl[3][1] - t0, tl[3][2] - t0);
  for (int kb = 0; kb < 19; ++kb)
    printf("kb %2d: producer issued %7lld | issuer saw full %7lld | issued+committed %7lld\n", kb, tl[0][kb] - t0, tl[1][kb] - t0, tl[2][kb] - t0);
  return 0;
}
