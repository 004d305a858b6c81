// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// A tiny single-OS-thread CUDA execution-model emulator so that the SIMT kernels
// under speech2affective_gestures_b200/csrc can be compiled with g++ and their
// *logic* (indexing, tiling, barriers, shuffles, atomics) checked on the CPU-only
// dev container before spending GPU minutes.  Each CUDA thread of a block runs as a
// ucontext fiber; __syncthreads()/warp shuffles are cooperative barriers that
// yield to a round-robin scheduler; blocks run one after the other.
//
// The product library (libs2ag_b200.so, built by nvcc for sm_100a) never includes
// this header: it is only reachable when S2AG_EMU is defined, which only
// tests/emu/build_emu.py does.
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __ldg(p) (*(p))

namespace emu {
struct State {
  dim3 tIdx, bIdx, bDim, gDim;
  ucontext_t main_ctx;
  std::vector<ucontext_t> ctx;
  std::vector<char*> stacks;
  std::vector<char> done;
  int cur = -1;
  int nthreads = 0;
  // block barrier
  int bar_count = 0;
  unsigned bar_gen = 0;
  // per-warp barriers + exchange slots
  std::vector<int> wbar_count;
  std::vector<unsigned> wbar_gen;
  std::vector<uint64_t> slots;
  std::function<void()> body;
  char* dyn_smem = nullptr;
};
inline State& S() { static State s; return s; }
static const size_t kStack = 256 * 1024;

inline void yield() {
  State& s = S();
  swapcontext(&s.ctx[s.cur], &s.main_ctx);
}
inline void trampoline() {
  State& s = S();
  s.body();
  s.done[s.cur] = 1;
  swapcontext(&s.ctx[s.cur], &s.main_ctx);
}
inline int lin_tid() {
  State& s = S();
  return s.tIdx.x + s.bDim.x * (s.tIdx.y + s.bDim.y * s.tIdx.z);
}
inline void block_barrier() {
  State& s = S();
  unsigned gen = s.bar_gen;
  if (++s.bar_count == s.nthreads) { s.bar_count = 0; s.bar_gen++; }
  else { while (s.bar_gen == gen) yield(); }
}
inline int warp_lanes(int w) { State& s = S(); return std::min(32, s.nthreads - w * 32); }
inline void warp_barrier() {
  State& s = S();
  int w = lin_tid() / 32;
  unsigned gen = s.wbar_gen[w];
  if (++s.wbar_count[w] == warp_lanes(w)) { s.wbar_count[w] = 0; s.wbar_gen[w]++; }
  else { while (s.wbar_gen[w] == gen) yield(); }
}
template <class T> inline T shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  State& s = S();
  int t = lin_tid(), w = t / 32;
  uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
  s.slots[t] = bits;
  warp_barrier();
  T r = v;
  if (src_lane >= 0 && src_lane < warp_lanes(w)) { uint64_t b = s.slots[w * 32 + src_lane]; memcpy(&r, &b, sizeof(T)); }
  warp_barrier();
  return r;
}
template <class F> inline void launch(dim3 grid, dim3 block, size_t smem, F&& f) {
  State& s = S();
  s.body = std::function<void()>(f);
  s.gDim = grid; s.bDim = block;
  s.nthreads = block.x * block.y * block.z;
  int nwarps = (s.nthreads + 31) / 32;
  if ((int)s.stacks.size() < s.nthreads) {
    size_t old = s.stacks.size();
    s.stacks.resize(s.nthreads);
    for (size_t i = old; i < s.stacks.size(); ++i) s.stacks[i] = (char*)malloc(kStack);
  }
  s.ctx.resize(s.nthreads); s.done.assign(s.nthreads, 0);
  s.wbar_count.assign(nwarps, 0); s.wbar_gen.assign(nwarps, 0); s.slots.assign(nwarps * 32, 0);
  std::vector<char> dyn(smem + 1024);
  s.dyn_smem = (char*)(((uintptr_t)dyn.data() + 1023) & ~(uintptr_t)1023);
  for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
    s.bIdx = dim3(bx, by, bz);
    s.bar_count = 0;
    std::fill(s.wbar_count.begin(), s.wbar_count.end(), 0);
    std::fill(s.done.begin(), s.done.end(), 0);
    for (int t = 0; t < s.nthreads; ++t) {
      getcontext(&s.ctx[t]);
      s.ctx[t].uc_stack.ss_sp = s.stacks[t];
      s.ctx[t].uc_stack.ss_size = kStack;
      s.ctx[t].uc_link = &s.main_ctx;
      makecontext(&s.ctx[t], (void (*)())trampoline, 0);
    }
    int remaining = s.nthreads;
    while (remaining > 0) {
      for (int t = 0; t < s.nthreads; ++t) {
        if (s.done[t]) continue;
        s.cur = t;
        s.tIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        swapcontext(&s.main_ctx, &s.ctx[t]);
        if (s.done[t]) --remaining;
      }
    }
  }
  s.dyn_smem = nullptr;
}
}  // namespace emu

#define threadIdx (emu::S().tIdx)
#define blockIdx (emu::S().bIdx)
#define blockDim (emu::S().bDim)
#define gridDim (emu::S().gDim)

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu::shfl_from(v, (emu::lin_tid() % 32) ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) { return emu::shfl_from(v, (emu::lin_tid() % 32) + d); }
template <class T> static inline T __shfl_sync(unsigned, T v, int l, int = 32) { return emu::shfl_from(v, l); }

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
#define __expf(x) expf(x)
#define __logf(x) logf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline void sincospif(float x, float* s, float* c) { *s = (float)sin(M_PI * (double)x); *c = (float)cos(M_PI * (double)x); }
static inline float cospif(float x) { return (float)cos(M_PI * (double)x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __uint2float_rn(unsigned x) { return (float)x; }
static inline float __int_as_float(int x) { float f; memcpy(&f, &x, 4); return f; }
static inline int __float_as_int(float x) { int i; memcpy(&i, &x, 4); return i; }

static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsyncD2D(void* d, const void* s, size_t n, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

#define S2AG_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::S().dyn_smem)
extern unsigned long long g_s2ag_launches;
#define S2AG_LAUNCH(kfn, grid, block, smem, stream, ...)                                  \
  do {                                                                                    \
    ++g_s2ag_launches;                                                                    \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kfn(__VA_ARGS__); });    \
  } while (0)
