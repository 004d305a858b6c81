// Development aid: clock64 marks of CTA 0 of the fused TCN block kernel (worker warp 0 and the MMA issuer).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -DS2AG_TCN_TIMELINE tools/tcn_timeline.cu -o tools/_build/tcn_timeline
#include <cstdarg>
#include <cstdlib>
#include <vector>
#include "../speech2affective_gestures_b200/csrc/umma_tcn.cu"
unsigned long long g_s2ag_launches = 0;
void s2ag_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
namespace s2ag { int g_engine = 0; namespace umma { int g_precision = 0; int g_dbg_flags = 0; } }

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 256, T = 34, C = 300, d = argc > 3 ? atoi(argv[3]) : 2;
  s2ag::tcnf::g_nstage_cap = argc > 2 ? atoi(argv[2]) : 0;
  const long n = (long)B * T * C;
  std::vector<float> h(n); for (long i = 0; i < n; ++i) h[i] = (float)(rand() % 2001 - 1000) * 1e-3f;
  std::vector<float> hv(2L * C * C); for (auto& v : hv) v = (float)(rand() % 2001 - 1000) * 1e-4f;
  std::vector<float> hg(C, 1.f), hb(C, 0.01f);
  float *x, *v1, *v2, *g1, *b1, *w1, *w2, *n1, *n2, *out, *ws;
  cudaMalloc(&x, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&v1, hv.size() * 4); cudaMalloc(&v2, hv.size() * 4);
  cudaMalloc(&g1, C * 4); cudaMalloc(&b1, C * 4); cudaMalloc(&w1, hv.size() * 4); cudaMalloc(&w2, hv.size() * 4);
  cudaMalloc(&n1, C * 4); cudaMalloc(&n2, C * 4);
  const long nws = s2ag_tcn_fused_ws_floats(T, C, d);
  cudaMalloc(&ws, nws * 4);
  cudaMemcpy(x, h.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(v1, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(v2, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(g1, hg.data(), C * 4, cudaMemcpyHostToDevice); cudaMemcpy(b1, hb.data(), C * 4, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 3; ++rep) {
    { long long z[3][16] = {}; cudaMemcpyToSymbol(s2ag::tcnf::g_tcn_tl, z, sizeof(z)); }
    int rc = s2ag_tcn_block_fused_fwd(x, v1, g1, b1, v2, g1, b1, w1, w2, n1, n2, nullptr, nullptr, out, ws, B, T, C, d, 0.f, 0, nullptr, 0);
    if (rc || cudaDeviceSynchronize() != cudaSuccess) { printf("failed rc=%d %s\n", rc, cudaGetErrorString(cudaGetLastError())); return 1; }
  }
  long long tl[3][16];
  cudaMemcpyFromSymbol(tl, s2ag::tcnf::g_tcn_tl, sizeof(tl));
  const long long t0 = tl[0][0];
  printf("d=%d ring slots cap %d: ", d, s2ag::tcnf::g_nstage_cap);
  printf("worker warp 0 (cycles since the start of x staging): staged %lld | acc1 ready %lld | y1 image written %lld | acc2(half 0) ready %lld | epilogue 2 done %lld\n",
         tl[0][1] - t0, tl[0][2] - t0, tl[0][3] - t0, tl[0][4] - t0, tl[0][5] - t0);
  printf("issuer: conv1 wait A %lld..%lld, half0 issued %lld, half1 issued %lld | conv2 wait A %lld..%lld, half0 issued %lld, half1 issued %lld\n",
         tl[1][0] - t0, tl[1][1] - t0, tl[1][2] - t0, tl[1][3] - t0, tl[1][4] - t0, tl[1][5] - t0, tl[1][6] - t0, tl[1][7] - t0);
  return 0;
}
