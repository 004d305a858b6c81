// Shared helpers for every kernel file of libs2ag_b200.so (sm_100a).
// Conventions of the C ABI (see include/s2ag.h): raw device pointers, explicit sizes,
// an opaque stream handle, int status return (0 = ok), no hidden allocation, no host sync.
#pragma once
#ifdef S2AG_EMU
#include "cuda_emu.h"   // tests/emu: CPU logic emulator, test infrastructure only
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#define S2AG_DYN_SMEM(type, name) extern __shared__ __align__(1024) unsigned char s2ag_dyn_smem_raw[]; \
  type* name = reinterpret_cast<type*>(s2ag_dyn_smem_raw)
extern unsigned long long g_s2ag_launches;
#define S2AG_LAUNCH(kfn, grid, block, smem, stream, ...)                                                  \
  do {                                                                                                    \
    ++g_s2ag_launches;                                                                                    \
    kfn<<<dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                \
  } while (0)
static inline cudaError_t cudaMemcpyAsyncD2D(void* d, const void* s, size_t n, cudaStream_t st) {
  return cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st);
}
#endif

#define S2AG_OK 0
#define S2AG_ERR_ARG (-1)
#define S2AG_ERR_LAUNCH (-2)
#define S2AG_ERR_UNSUPPORTED (-3)

void s2ag_set_error(const char* fmt, ...);

#define S2AG_CHECK_ARG(cond)                                                        \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      s2ag_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond);     \
      return S2AG_ERR_ARG;                                                          \
    }                                                                               \
  } while (0)

#define S2AG_CHECK_LAUNCH()                                                         \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      s2ag_set_error("%s:%d: launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return S2AG_ERR_LAUNCH;                                                       \
    }                                                                               \
  } while (0)

// device-side assertion (mirrors the device assert nn.Embedding raises on an out-of-range index)
#ifdef S2AG_EMU
#define S2AG_DEVICE_TRAP() abort()
#else
#define S2AG_DEVICE_TRAP() __trap()
#endif

static inline int s2ag_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

#ifndef S2AG_EMU
// SMs of the current device (queried once; 148 on B200): the residency bound of the persistent kernels and the grid
// multiple of the grid-stride ones
static inline int s2ag_sm_count() {
  static int sms = 0;
  if (sms <= 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}
#else
static inline int s2ag_sm_count() { return 148; }
#endif

// activation codes shared by every epilogue (mirrors the reference's nn.LeakyReLU slopes,
// SURVEY Appendix B): 0 none, 1 ReLU, 2 LeakyReLU(slope)
#define S2AG_ACT_NONE 0
#define S2AG_ACT_RELU 1
#define S2AG_ACT_LEAKY 2

__device__ __forceinline__ float s2ag_act(float v, int act, float slope) {
  if (act == S2AG_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == S2AG_ACT_LEAKY) return v > 0.f ? v : v * slope;
  return v;
}
// derivative of the activation expressed through its OUTPUT (valid for slope >= 0):
__device__ __forceinline__ float s2ag_act_grad_from_out(float out, int act, float slope) {
  if (act == S2AG_ACT_RELU) return out > 0.f ? 1.f : 0.f;
  if (act == S2AG_ACT_LEAKY) return out > 0.f ? 1.f : slope;
  return 1.f;
}

// Counter-based dropout RNG (stateless: the backward pass regenerates the same mask from
// (seed, element index)).  One 32-bit mix per element; keep = u >= p * 2^32.
__device__ __forceinline__ unsigned s2ag_hash32(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (unsigned)(z >> 32);
}
__device__ __forceinline__ float s2ag_dropout_scale(unsigned long long seed, unsigned long long idx, float p) {
  // returns 0 (dropped) or 1/(1-p) (kept); p == 0 -> 1
  if (p <= 0.f) return 1.f;
  unsigned u = s2ag_hash32(seed, idx);
  unsigned thr = (unsigned)(p * 4294967296.0);
  return u >= thr ? 1.f / (1.f - p) : 0.f;
}

__device__ __forceinline__ float s2ag_warp_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ double s2ag_warp_sum_d(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ float s2ag_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
