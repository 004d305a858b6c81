// Bring-up / regression check of the tcgen05 GEMM engine (gemm_umma.cuh) against the exact-fp32
// SIMT engine and a sampled double-precision CPU reference.  Standalone: nvcc, no torch.
//   tools/_build/umma_selftest [dbg_flags] [precision]
#include "../speech2affective_gestures_b200/csrc/gemm.cuh"
#include <cstdarg>
#include <vector>
#include <random>
#include <string>

unsigned long long g_s2ag_launches = 0;
void s2ag_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
namespace s2ag { int g_engine = 0; namespace umma { int g_precision = 0; int g_dbg_flags = 0; } }
using namespace s2ag;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static std::vector<float> rnd(size_t n, unsigned seed, float scale = 1.f) {
  std::mt19937 g(seed); std::uniform_real_distribution<float> d(-scale, scale);
  std::vector<float> v(n); for (auto& x : v) x = d(g); return v;
}
static float* dev(const std::vector<float>& h) { float* p; CK(cudaMalloc(&p, h.size() * 4 + 64)); CK(cudaMemcpy(p, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); return p; }

struct Result { double max_vs_simt, max_ref, max_vs_cpu; float ms_umma, ms_simt; };

template <class LdA, class LdB, class HostA, class HostB>
static Result run_case(const char* name, LdA a, LdB b, HostA ha, HostB hb, int M, int N, int K, int nbatch, int splitk,
                       bool time_it) {
  const size_t nc = (size_t)nbatch * M * N;
  float *c_u, *c_s;
  CK(cudaMalloc(&c_u, nc * 4)); CK(cudaMalloc(&c_s, nc * 4));
  CK(cudaMemset(c_u, 0, nc * 4)); CK(cudaMemset(c_s, 0, nc * 4));
  EpiGeneric eu = make_epi(c_u, (long)N, nullptr, 0, 0.f, splitk > 1 ? 2 : 0); eu.bstride = (long)M * N;
  EpiGeneric es = make_epi(c_s, (long)N, nullptr, 0, 0.f, splitk > 1 ? 2 : 0); es.bstride = (long)M * N;
  umma::launch(a, b, eu, M, N, K, nbatch, splitk, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-28s UMMA kernel failed: %s\n", name, cudaGetErrorString(e)); exit(3); }
  launch_gemm_simt(a, b, es, M, N, K, nbatch, splitk, nullptr);
  CK(cudaDeviceSynchronize());
  std::vector<float> hu(nc), hs(nc);
  CK(cudaMemcpy(hu.data(), c_u, nc * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hs.data(), c_s, nc * 4, cudaMemcpyDeviceToHost));
  Result r{0, 0, 0, 0, 0};
  for (size_t i = 0; i < nc; ++i) { r.max_vs_simt = fmax(r.max_vs_simt, fabs((double)hu[i] - hs[i])); r.max_ref = fmax(r.max_ref, fabs((double)hs[i])); }
  std::mt19937 g(7);
  for (int s = 0; s < 400; ++s) {
    int bb = g() % nbatch, m = g() % M, n = g() % N;
    double acc = 0; for (int k = 0; k < K; ++k) acc += (double)ha(bb, m, k) * (double)hb(bb, n, k);
    r.max_vs_cpu = fmax(r.max_vs_cpu, fabs(acc - hu[((size_t)bb * M + m) * N + n]));
  }
  if (time_it) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) umma::launch(a, b, eu, M, N, K, nbatch, splitk, nullptr);
    cudaEventRecord(e0); for (int w = 0; w < 10; ++w) umma::launch(a, b, eu, M, N, K, nbatch, splitk, nullptr); cudaEventRecord(e1);
    CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&r.ms_umma, e0, e1); r.ms_umma /= 10;
    for (int w = 0; w < 2; ++w) launch_gemm_simt(a, b, es, M, N, K, nbatch, splitk, nullptr);
    cudaEventRecord(e0); for (int w = 0; w < 5; ++w) launch_gemm_simt(a, b, es, M, N, K, nbatch, splitk, nullptr); cudaEventRecord(e1);
    CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&r.ms_simt, e0, e1); r.ms_simt /= 5;
  }
  double fl = 2.0 * M * N * K * nbatch;
  printf("%-28s M=%d N=%d K=%d nb=%d sk=%d | rel(umma-simt)=%.3e rel(umma-cpu64)=%.3e", name, M, N, K, nbatch, splitk,
         r.max_vs_simt / r.max_ref, r.max_vs_cpu / r.max_ref);
  if (time_it) printf(" | umma %.3f ms (%.1f TF/s) simt %.3f ms (%.1f TF/s)", r.ms_umma, fl / r.ms_umma * 1e-9, r.ms_simt, fl / r.ms_simt * 1e-9);
  printf("\n"); fflush(stdout);
  cudaFree(c_u); cudaFree(c_s);
  return r;
}

int main(int argc, char** argv) {
  umma::g_dbg_flags = argc > 1 ? atoi(argv[1]) : 0;
  umma::g_precision = argc > 2 ? atoi(argv[2]) : 0;
  printf("umma selftest: dbg_flags=%d precision=%d\n", umma::g_dbg_flags, umma::g_precision);
  {  // tiny, exact-in-bf16 operands: any layout/descriptor error shows as O(1) error
    int M = 128, N = 64, K = 32;
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[(size_t)m * K + k] = (float)((m * 7 + k * 3) % 11 - 5);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[(size_t)n * K + k] = (float)((n * 5 + k * 2) % 7 - 3);
    float *dA = dev(A), *dB = dev(B);
    run_case("tiny exact", LdPlain<true>{dA, (long)K, 1, 0}, LdPlain<true>{dB, (long)K, 1, 0},
             [&](int, int m, int k) { return A[(size_t)m * K + k]; }, [&](int, int n, int k) { return B[(size_t)n * K + k]; },
             M, N, K, 1, 1, false);
  }
  {  // odd shapes: all tails
    int M = 129, N = 130, K = 45;
    auto A = rnd((size_t)M * K, 1), B = rnd((size_t)N * K, 2);
    float *dA = dev(A), *dB = dev(B);
    run_case("tails", LdPlain<true>{dA, (long)K, 1, 0}, LdPlain<true>{dB, (long)K, 1, 0},
             [&](int, int m, int k) { return A[(size_t)m * K + k]; }, [&](int, int n, int k) { return B[(size_t)n * K + k]; },
             M, N, K, 1, 1, false);
  }
  {  // GRU input projection, both directions batched (B=256): x[8704,600] @ W[2][900,600]^T
    int M = 8704, N = 900, K = 600;
    auto A = rnd((size_t)M * K, 3), B = rnd((size_t)2 * N * K, 4, 0.06f);
    float *dA = dev(A), *dB = dev(B);
    run_case("gru input projection", LdPlain<true>{dA, (long)K, 1, 0}, LdPlain<true>{dB, (long)K, 1, (long)N * K},
             [&](int, int m, int k) { return A[(size_t)m * K + k]; },
             [&](int bb, int n, int k) { return B[((size_t)bb * N + n) * K + k]; }, M, N, K, 2, 1, true);
  }
  {  // causal TCN conv as GEMM: rows (b,t), K = (tap, c), dilation 4
    int Bt = 256, T = 34, C = 300, d = 4, M = Bt * T, K = 2 * C;
    auto X = rnd((size_t)M * C, 5), W = rnd((size_t)C * K, 6, 0.05f);
    float *dX = dev(X), *dW = dev(W);
    LdConv<ORDER_KKC> la{dX, T, 1, C, T, 1, 2, 1, 1, 1, d, 1, +1, -d, 0, (long)C};
    run_case("tcn conv (dilation 4)", la, LdPlain<true>{dW, (long)K, 1, 0},
             [&](int, int m, int k) { int t = m % T, b = m / T, j = k / C, c = k % C; int ts = t + (j - 1) * d;
               return ts < 0 ? 0.f : X[((size_t)b * T + ts) * C + c]; },
             [&](int, int n, int k) { return W[(size_t)n * K + k]; }, M, C, K, 1, 1, true);
  }
  {  // weight gradient: dW[900,600] += dgi^T x, contraction over 8704 rows, split-K
    int R = 8704, No = 900, Ki = 600;
    auto G = rnd((size_t)R * No, 7), X = rnd((size_t)R * Ki, 8);
    float *dG = dev(G), *dX = dev(X);
    int sk = pick_splitk(No, Ki, R, 1);
    run_case("wgrad split-K", LdPlain<false>{dG, 1, (long)No, 0}, LdPlain<false>{dX, 1, (long)Ki, 0},
             [&](int, int m, int k) { return G[(size_t)k * No + m]; }, [&](int, int n, int k) { return X[(size_t)k * Ki + n]; },
             No, Ki, R, 1, sk, true);
  }
  {  // small-N head: [8704,300] @ [32,300]^T
    int M = 8704, N = 32, K = 300;
    auto A = rnd((size_t)M * K, 9), B = rnd((size_t)N * K, 10);
    float *dA = dev(A), *dB = dev(B);
    run_case("linear 300->32", LdPlain<true>{dA, (long)K, 1, 0}, LdPlain<true>{dB, (long)K, 1, 0},
             [&](int, int m, int k) { return A[(size_t)m * K + k]; }, [&](int, int n, int k) { return B[(size_t)n * K + k]; },
             M, N, K, 1, 1, true);
  }
  {  // TCN data-gradient: A = gathered dY (sgn=-1), B = weight view W[co][tap][ci] read along co (LdWdgrad)
    int Bt = 3, T = 34, C = 300, d = 2, M = Bt * T, K = 2 * C;
    auto G = rnd((size_t)M * C, 11), W = rnd((size_t)C * K, 12, 0.05f);
    float *dG = dev(G), *dW = dev(W);
    LdConv<ORDER_KKC> la{dG, T, 1, C, T, 1, 2, 1, 1, 1, d, 1, -1, d, 0, (long)C};
    LdWdgrad<ORDER_KKC> lb{dW, C, 2, (long)K, 1, (long)C};
    run_case("tcn dgrad", la, lb,
             [&](int, int m, int k) { int t = m % T, b = m / T, j = k / C, c = k % C; int ts = t + (1 - j) * d;
               return (ts < 0 || ts >= T) ? 0.f : G[((size_t)b * T + ts) * C + c]; },
             [&](int, int n, int k) { int kk = k / C, co = k % C; return W[(size_t)co * K + kk * C + n]; }, M, C, K, 1, 1, false);
  }
  {  // TCN weight-gradient: rows co, cols (tap,c), contraction over pixels
    int Bt = 3, T = 34, C = 300, d = 2, M = Bt * T, K = 2 * C;
    auto G = rnd((size_t)M * C, 13), X = rnd((size_t)M * C, 14);
    float *dG = dev(G), *dX = dev(X);
    LdConv<ORDER_KKC> lx{dX, T, 1, C, T, 1, 2, 1, 1, 1, d, 1, +1, -d, 0, (long)C};
    int sk = pick_splitk(C, K, M, 1);
    run_case("tcn wgrad", LdPlain<false>{dG, 1, (long)C, 0}, LdT<LdConv<ORDER_KKC>>{lx},
             [&](int, int m, int k) { return G[(size_t)k * C + m]; },
             [&](int, int n, int k) { int t = k % T, b = k / T, j = n / C, c = n % C; int ts = t + (j - 1) * d;
               return ts < 0 ? 0.f : X[((size_t)b * T + ts) * C + c]; }, C, K, M, 1, sk, false);
  }
  {  // unaligned rows (K = 150) and N = 27 (-> BN 32)
    int M = 8704, N = 27, K = 150;
    auto A = rnd((size_t)M * K, 15), B = rnd((size_t)N * K, 16);
    float *dA = dev(A), *dB = dev(B);
    run_case("linear 150->27 (unaligned)", LdPlain<true>{dA, (long)K, 1, 0}, LdPlain<true>{dB, (long)K, 1, 0},
             [&](int, int m, int k) { return A[(size_t)m * K + k]; }, [&](int, int n, int k) { return B[(size_t)n * K + k]; },
             M, N, K, 1, 1, true);
  }
  {  // Conv1d implicit GEMM, PyTorch weight order (c, k): MFCC conv1 71->64, k5, pad 2, L=37
    int Nb = 64, L = 37, Cin = 71, Cout = 64, KH = 5, M = Nb * L, K = Cin * KH;
    auto X = rnd((size_t)M * Cin, 17), W = rnd((size_t)Cout * K, 18, 0.1f);
    float *dX = dev(X), *dW = dev(W);
    LdConv<ORDER_CKK> la{dX, L, 1, Cin, L, 1, KH, 1, 1, 1, 1, 1, +1, -2, 0, (long)Cin};
    run_case("conv1d 71->64 k5", la, LdPlain<true>{dW, (long)K, 1, 0},
             [&](int, int m, int k) { int l = m % L, b = m / L, c = k / KH, j = k % KH; int ls = l + j - 2;
               return (ls < 0 || ls >= L) ? 0.f : X[((size_t)b * L + ls) * Cin + c]; },
             [&](int, int n, int k) { return W[(size_t)n * K + k]; }, M, Cout, K, 1, 1, true);
  }
  printf("selftest done\n");
  return 0;
}
