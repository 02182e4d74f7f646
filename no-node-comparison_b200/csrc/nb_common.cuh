// nb_common.cuh — shared device helpers for the EGNO / SEGNO kernels (sm_100a).
#pragma once

#ifndef NB_EMU
#include <cuda_runtime.h>
#define NB_DYN_SMEM(name) extern __shared__ __align__(16) float name[]
// Programmatic dependent launch, opt-in (NB_B200_PDL=1).  Measured on B200, cfg3 under graph replay: 3.621 ms per step
// without, 3.631 ms with — with the wait at the very top of every kernel only the launch latency can overlap, and a
// graph replay has already removed it; the big kernels fill every SM's shared memory, so a dependent CTA cannot become
// resident before a primary CTA exits anyway.  Kept as a switch because it costs nothing when off (the two instructions
// are no-ops without a programmatic edge).  When on:
// every kernel of the library is launched with programmatic stream serialization allowed and begins with NB_PDL_ENTER():
// `griddepcontrol.launch_dependents` lets the NEXT kernel of the stream be scheduled as soon as every CTA of this one has
// started (its CTAs become resident as SM resources free up instead of after the last CTA has drained and the launch
// latency has elapsed), `griddepcontrol.wait` then blocks until the PREVIOUS kernel has completed and flushed its
// memory, before anything is read or written.  Nothing precedes the wait, so the memory model is the plain stream order;
// what is gained is the launch latency and the ramp between ~80 kernels per training step (also inside a captured graph,
// where the dependency becomes a programmatic edge).
#define NB_PDL_ENTER() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
int nb_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline void nb_launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = nb_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#define NB_LAUNCH(kern, grid, block, smem, stream, ...) nb_launch_kernel(kern, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
#else
#define NB_PDL_ENTER()
#endif

#include <stdint.h>

#include "../../include/nbody_b200.h"

#define NB_H 64            // hidden width (== NB_HIDDEN)
#define NB_LDA 68          // row stride (floats) of activation tiles in shared memory: 64 + 4 pad
#define NB_TILE 128        // rows per tile
#define NB_THREADS 256     // threads per CTA of the tiled kernels
#define NB_MAX_EF NB_MAX_EDGE_FEA

// ----------------------------------------------------------------------------- activations
// SiLU and its derivative from one sigmoid evaluation (ex2.approx + rcp.approx: ~2 ulp).
#ifndef NB_EMU
// sigmoid(x) = 1 / (1 + 2^(-x log2 e)) with the two MUFU approximations issued directly (no range fix-up code:
// ex2 saturates to 0 / +inf and rcp(+inf) = 0, which are the correct limits)
__device__ __forceinline__ float nb_sigmoid(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
  return r;
}
// The same sigmoid with the reciprocal on the FMA pipe (no second MUFU operation): bit-trick seed (10 % off) and three
// Newton steps r <- r + r (1 - y r), each halving the exponent of the error: 1e-1 -> 1e-2 -> 1e-4 -> 6e-8 (fp32 rounding;
// k_tc_selftest mode 11 measures it against rcp.approx).  The exponent argument is clamped so that y = 1 + 2^a stays
// finite (a <= 126: sigmoid = 2^-126, SiLU -> 0, the correct limit).  The edge forward kernel is bound by the 16
// MUFU lanes per SM (two MUFU operations per SiLU, 24 576 SiLUs per 128-edge tile); this variant halves that load.
__device__ __forceinline__ float nb_sigmoid_fma(float x) {
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fminf(x * -1.4426950408889634f, 126.0f)));
  const float y = 1.0f + t;
  float r = __uint_as_float(0x7EF311C7u - __float_as_uint(y));
#pragma unroll
  for (int it = 0; it < 3; ++it) r = fmaf(r, fmaf(-y, r, 1.0f), r);
  return r;
}
#else
__device__ __forceinline__ float nb_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float nb_sigmoid_fma(float x) { return nb_sigmoid(x); }
#endif
__device__ __forceinline__ float nb_silu_fma(float x) { return x * nb_sigmoid_fma(x); }
__device__ __forceinline__ float nb_silu(float x) { return x * nb_sigmoid(x); }
__device__ __forceinline__ void nb_silu_grad(float x, float& y, float& dy) {
  float s = nb_sigmoid(x);
  y = x * s;
  dy = fmaf(y, 1.0f - s, s);  // s (1 + x (1 - s))
}
__device__ __forceinline__ float nb_dsilu(float x) {
  float s = nb_sigmoid(x);
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float4 nb_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void nb_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ----------------------------------------------------------------------------- tile GEMM (SIMT fp32)
// acc[q][c] += sum_k A[(ty + 16 q)][k] * B[k][4 tx + c],  k < 64.
//   A: shared, row-major, stride NB_LDA.   B: shared, k-major [64][64].
// 256 threads as 16 (ty) x 16 (tx); each thread owns RPT interleaved rows x 4 consecutive columns.
// Interleaving rows across ty keeps the A reads of one warp (two ty values) on distinct banks;
// the 16 tx lanes read one contiguous 256-byte B row per k.
template <int RPT>
__device__ __forceinline__ void nb_tile_gemm(const float* __restrict__ As, const float* __restrict__ Bs,
                                             float (&acc)[RPT][4], int ty, int tx) {
#pragma unroll 4
  for (int k0 = 0; k0 < NB_H; k0 += 4) {
    float4 b0 = nb_ld4(Bs + (k0 + 0) * NB_H + tx * 4);
    float4 b1 = nb_ld4(Bs + (k0 + 1) * NB_H + tx * 4);
    float4 b2 = nb_ld4(Bs + (k0 + 2) * NB_H + tx * 4);
    float4 b3 = nb_ld4(Bs + (k0 + 3) * NB_H + tx * 4);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      float4 a = nb_ld4(As + (ty + 16 * q) * NB_LDA + k0);
      acc[q][0] = fmaf(a.x, b0.x, acc[q][0]);
      acc[q][1] = fmaf(a.x, b0.y, acc[q][1]);
      acc[q][2] = fmaf(a.x, b0.z, acc[q][2]);
      acc[q][3] = fmaf(a.x, b0.w, acc[q][3]);
      acc[q][0] = fmaf(a.y, b1.x, acc[q][0]);
      acc[q][1] = fmaf(a.y, b1.y, acc[q][1]);
      acc[q][2] = fmaf(a.y, b1.z, acc[q][2]);
      acc[q][3] = fmaf(a.y, b1.w, acc[q][3]);
      acc[q][0] = fmaf(a.z, b2.x, acc[q][0]);
      acc[q][1] = fmaf(a.z, b2.y, acc[q][1]);
      acc[q][2] = fmaf(a.z, b2.z, acc[q][2]);
      acc[q][3] = fmaf(a.z, b2.w, acc[q][3]);
      acc[q][0] = fmaf(a.w, b3.x, acc[q][0]);
      acc[q][1] = fmaf(a.w, b3.y, acc[q][1]);
      acc[q][2] = fmaf(a.w, b3.z, acc[q][2]);
      acc[q][3] = fmaf(a.w, b3.w, acc[q][3]);
    }
  }
}

// Weight-gradient tile: acc[i][j] += scale * sum_{r < nrows} G[r][4 wo + i] * A[r][4 wk + j]
//   G, A: shared, row-major, stride NB_LDA.  256 threads as 16 (wo) x 16 (wk): a 64 x 64 result.
__device__ __forceinline__ void nb_tile_wgrad(const float* __restrict__ Gs, const float* __restrict__ As, int nrows,
                                              float (&acc)[4][4], int wo, int wk) {
#pragma unroll 2
  for (int r = 0; r < nrows; ++r) {
    float4 g = nb_ld4(Gs + r * NB_LDA + wo * 4);
    float4 a = nb_ld4(As + r * NB_LDA + wk * 4);
    acc[0][0] = fmaf(g.x, a.x, acc[0][0]);
    acc[0][1] = fmaf(g.x, a.y, acc[0][1]);
    acc[0][2] = fmaf(g.x, a.z, acc[0][2]);
    acc[0][3] = fmaf(g.x, a.w, acc[0][3]);
    acc[1][0] = fmaf(g.y, a.x, acc[1][0]);
    acc[1][1] = fmaf(g.y, a.y, acc[1][1]);
    acc[1][2] = fmaf(g.y, a.z, acc[1][2]);
    acc[1][3] = fmaf(g.y, a.w, acc[1][3]);
    acc[2][0] = fmaf(g.z, a.x, acc[2][0]);
    acc[2][1] = fmaf(g.z, a.y, acc[2][1]);
    acc[2][2] = fmaf(g.z, a.z, acc[2][2]);
    acc[2][3] = fmaf(g.z, a.w, acc[2][3]);
    acc[3][0] = fmaf(g.w, a.x, acc[3][0]);
    acc[3][1] = fmaf(g.w, a.y, acc[3][1]);
    acc[3][2] = fmaf(g.w, a.z, acc[3][2]);
    acc[3][3] = fmaf(g.w, a.w, acc[3][3]);
  }
}

// Stage a 64x64 operand into shared memory as Bs[k][n] = src[k * sk + n * sn] * scale.
// kmax: number of valid k (rows of the operand); rows k >= kmax are zero (operands narrower than 64)
__device__ __forceinline__ void nb_stage_b(float* __restrict__ Bs, const float* __restrict__ src, int64_t sk,
                                           int64_t sn, float scale, int tid, int kmax = NB_H) {
  if (sn == 1) {  // rows of src are contiguous in n: coalesced reads, conflict-free writes
    for (int idx = tid; idx < NB_H * NB_H; idx += NB_THREADS) {
      int k = idx >> 6, n = idx & 63;
      Bs[idx] = k < kmax ? __ldg(src + (int64_t)k * sk + n) * scale : 0.f;
    }
  } else {        // read along k (contiguous when sk == 1), transposing on the way in
    for (int idx = tid; idx < NB_H * NB_H; idx += NB_THREADS) {
      int n = idx >> 6, k = idx & 63;
      Bs[k * NB_H + n] = k < kmax ? __ldg(src + (int64_t)k * sk + (int64_t)n * sn) * scale : 0.f;
    }
  }
}

// Sum of a per-thread value over the 16 tx lanes that share a row (lanes of one half-warp).
__device__ __forceinline__ float nb_reduce_tx(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// ----------------------------------------------------------------------------- host-side helpers
#ifndef NB_EMU
#define NB_SET_SMEM(kern, bytes)                                                                     \
  do {                                                                                                \
    if ((bytes) > 48 * 1024)                                                                          \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));          \
  } while (0)
#else
#define NB_SET_SMEM(kern, bytes) \
  do {                           \
  } while (0)
#endif

void nb_set_error(const char* fmt, ...);
int nb_check_launch(const char* what);
int nb_num_sms();
