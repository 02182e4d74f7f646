// cuda_emu.h — a tiny single-threaded CUDA *emulator* for host-side unit tests.
//
// TEST INFRASTRUCTURE ONLY.  It lets the CPU-only build container compile the very same
// kernel sources (csrc/*.cu) with g++ and run them block by block, thread by thread, so
// indexing / reduction / barrier logic can be checked against the oracle before spending
// GPU minutes.  The product library is built by nvcc for sm_100a and never includes this
// file; nothing in the Python package can load the emulated library.
//
// Model: each CUDA thread of a block is a ucontext fiber; __syncthreads() and the warp
// shuffles yield to a round-robin scheduler until every (live) thread of the block / warp
// has arrived.  Blocks run one after another.  Only the subset of CUDA used by csrc/ is
// provided.
#pragma once
#ifndef NB_EMU
#error "cuda_emu.h is only for -DNB_EMU host builds"
#endif

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

// ------------------------------------------------------------------ qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ------------------------------------------------------------------ vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

// ------------------------------------------------------------------ runtime shims
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memmove(d, s, n);
  return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ------------------------------------------------------------------ the scheduler
namespace nbemu {

struct State {
  ucontext_t main_ctx;
  std::vector<ucontext_t> ctx;
  std::vector<char*> stacks;
  std::vector<char> done;
  int nthreads = 0, nlive = 0, cur = 0;
  // block barrier
  int bar_count = 0;
  unsigned bar_gen = 0;
  // warp barriers + shuffle slots
  std::vector<int> wbar_count;
  std::vector<unsigned> wbar_gen;
  std::vector<uint32_t> slots;  // [nthreads]
  std::function<void()> body;
  unsigned char* dyn_smem = nullptr;
};

extern State g;
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;

static inline void yield() { swapcontext(&g.ctx[g.cur], &g.main_ctx); }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

static inline void syncthreads() {
  unsigned gen = g.bar_gen;
  if (++g.bar_count >= g.nlive) {
    g.bar_count = 0;
    g.bar_gen++;
    return;
  }
  while (g.bar_gen == gen) yield();
}

static inline void syncwarp_full() {
  int w = g.cur >> 5;
  int lanes = g.nthreads - (w << 5);
  if (lanes > 32) lanes = 32;
  unsigned gen = g.wbar_gen[w];
  if (++g.wbar_count[w] >= lanes) {
    g.wbar_count[w] = 0;
    g.wbar_gen[w]++;
    return;
  }
  while (g.wbar_gen[w] == gen) yield();
}

static inline uint32_t shfl_idx(uint32_t v, int src_lane) {
  int w = g.cur >> 5;
  g.slots[g.cur] = v;
  syncwarp_full();
  uint32_t r = g.slots[(w << 5) + (src_lane & 31)];
  syncwarp_full();
  return r;
}

}  // namespace nbemu

#define threadIdx (nbemu::g_threadIdx)
#define blockIdx (nbemu::g_blockIdx)
#define blockDim (nbemu::g_blockDim)
#define gridDim (nbemu::g_gridDim)

static inline void __syncthreads() { nbemu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { nbemu::syncwarp_full(); }

static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = nbemu::shfl_idx(u, (nbemu::g.cur & 31) ^ lane_mask);
  memcpy(&v, &u, 4);
  return v;
}
static inline float __shfl_down_sync(unsigned, float v, int delta) {
  uint32_t u;
  memcpy(&u, &v, 4);
  int lane = nbemu::g.cur & 31;
  int src = lane + delta > 31 ? lane : lane + delta;
  u = nbemu::shfl_idx(u, src);
  memcpy(&v, &u, 4);
  return v;
}
static inline float __shfl_sync(unsigned, float v, int src) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = nbemu::shfl_idx(u, src);
  memcpy(&v, &u, 4);
  return v;
}
static inline int __shfl_sync(unsigned, int v, int src) { return (int)nbemu::shfl_idx((uint32_t)v, src); }

// ------------------------------------------------------------------ intrinsics
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __ldg(const float* p) { return *p; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline float2 __ldg(const float2* p) { return *p; }
static inline int64_t __ldg(const int64_t* p) { return *p; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline int atomicCAS(int* p, int cmp, int val) { int o = *p; if (o == cmp) *p = val; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline float fmaf_emu(float a, float b, float c) { return fmaf(a, b, c); }
#ifndef __sincosf
static inline void sincosf_emu(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
#endif

#define NB_DYN_SMEM(name) float* name = reinterpret_cast<float*>(nbemu::g.dyn_smem)
#define NB_LAUNCH(kern, grid, block, smem, stream, ...) \
  nbemu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
