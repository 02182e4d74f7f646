// nb_edge_tc.cuh — the fused E_GCL edge tile on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same math and same work decomposition as nb_edge.cuh (units of graph-instances, 128-edge tiles, receiver
// reductions over contiguous rows, no atomics); the two 64x64 contractions per edge (phi_e layer 2, phi_x layer 1)
// and, in the backward, the two data-gradient and two weight-gradient contractions run as tcgen05.mma with
// split-bf16 operands and fp32 accumulators in TMEM (nb_tc.cuh).
//
// Thread mapping: 256 threads per 128-row tile.  Thread (warp w, lane l) owns row 32 (w & 3) + l — the TMEM lane
// quarter its warp may access — and the 32 columns of half (w >> 2).  Everything per-row lives in registers; only
// the cross-row reductions (M_i, Fsum_i, gP, gQ, gx) go through shared memory.
#pragma once
#ifndef NB_EMU
#include "nb_edge.cuh"
#include "nb_tc.cuh"

// row geometry held in registers by the row's two threads
struct NbRowRegs {
  int ni, nj;      // receiver / sender node (global)
  float dx, dy, dz, r2;
  float e[NB_MAX_EF];
  bool valid;
};

__device__ __forceinline__ NbRowRegs nb_row_regs(const NbEdgeGeom& g, const float* __restrict__ x,
                                                 const float* __restrict__ ef, int gt0, int r0, int nv, int row) {
  NbRowRegs R;
  R.valid = row < nv;
  R.dx = R.dy = R.dz = R.r2 = 0.f;
#pragma unroll
  for (int f = 0; f < NB_MAX_EF; ++f) R.e[f] = 0.f;
  if (R.valid) {
    int r = r0 + row;
    int lg = r / g.EPG, rem = r - lg * g.EPG;
    int i = rem / (g.N - 1), jj = rem - i * (g.N - 1);
    int j = jj + (jj >= i ? 1 : 0);
    int gt = gt0 + lg;
    R.ni = gt * g.N + i;
    R.nj = gt * g.N + j;
    R.dx = __ldg(x + (int64_t)R.ni * 3 + 0) - __ldg(x + (int64_t)R.nj * 3 + 0);
    R.dy = __ldg(x + (int64_t)R.ni * 3 + 1) - __ldg(x + (int64_t)R.nj * 3 + 1);
    R.dz = __ldg(x + (int64_t)R.ni * 3 + 2) - __ldg(x + (int64_t)R.nj * 3 + 2);
    R.r2 = R.dx * R.dx + R.dy * R.dy + R.dz * R.dz;
    int64_t eoff = ((int64_t)(gt % g.B) * g.EPG + rem) * g.nef;
#pragma unroll
    for (int f = 0; f < NB_MAX_EF; ++f)
      if (f < g.nef) R.e[f] = __ldg(ef + eoff + f);
  } else {
    R.ni = R.nj = gt0 * g.N;
  }
  return R;
}

// pre-activation of the first edge layer, 8 consecutive columns starting at c0
__device__ __forceinline__ void nb_pre1_8(const float* __restrict__ P, const float* __restrict__ Q, const NbRowRegs& R,
                                          int c0, const float* __restrict__ vwr, const float* __restrict__ vwe, int nef,
                                          float (&v)[8]) {
  float4 p0 = nb_ld4(P + (int64_t)R.ni * NB_H + c0), p1 = nb_ld4(P + (int64_t)R.ni * NB_H + c0 + 4);
  float4 q0 = nb_ld4(Q + (int64_t)R.nj * NB_H + c0), q1 = nb_ld4(Q + (int64_t)R.nj * NB_H + c0 + 4);
  float4 w0 = nb_ld4(vwr + c0), w1 = nb_ld4(vwr + c0 + 4);
  v[0] = fmaf(w0.x, R.r2, p0.x + q0.x);
  v[1] = fmaf(w0.y, R.r2, p0.y + q0.y);
  v[2] = fmaf(w0.z, R.r2, p0.z + q0.z);
  v[3] = fmaf(w0.w, R.r2, p0.w + q0.w);
  v[4] = fmaf(w1.x, R.r2, p1.x + q1.x);
  v[5] = fmaf(w1.y, R.r2, p1.y + q1.y);
  v[6] = fmaf(w1.z, R.r2, p1.z + q1.z);
  v[7] = fmaf(w1.w, R.r2, p1.w + q1.w);
#pragma unroll
  for (int f = 0; f < NB_MAX_EF; ++f)
    if (f < nef) {
      float4 e0 = nb_ld4(vwe + f * NB_H + c0), e1 = nb_ld4(vwe + f * NB_H + c0 + 4);
      v[0] = fmaf(e0.x, R.e[f], v[0]);
      v[1] = fmaf(e0.y, R.e[f], v[1]);
      v[2] = fmaf(e0.z, R.e[f], v[2]);
      v[3] = fmaf(e0.w, R.e[f], v[3]);
      v[4] = fmaf(e1.x, R.e[f], v[4]);
      v[5] = fmaf(e1.y, R.e[f], v[5]);
      v[6] = fmaf(e1.z, R.e[f], v[6]);
      v[7] = fmaf(e1.w, R.e[f], v[7]);
    }
}

// stage a [64][64] fp32 weight matrix (row-major, row = output unit) as a split-bf16 operand tile
__device__ __forceinline__ void nb_tc_stage_weight(unsigned char* hi, unsigned char* lo, const float* __restrict__ W,
                                                   int tid) {
  for (int idx = tid; idx < 64 * 8; idx += NB_THREADS) {
    int o = idx >> 3, j = idx & 7;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(W + o * NB_H + 8 * j + i);
    nb_tc_store8(hi, lo, o, j, v);
  }
}

// value of element (r, column pair cp) of a split-bf16 tile: returns columns 2cp, 2cp+1 as hi + lo
__device__ __forceinline__ float2 nb_tc_load_pair(const unsigned char* hi, const unsigned char* lo, int r, int cp) {
  uint32_t off = (uint32_t)r * NB_TC_ROW_BYTES + (uint32_t)((((cp >> 2) ^ (r & 7))) << 4) + (uint32_t)(cp & 3) * 4;
  uint32_t h = *reinterpret_cast<const uint32_t*>(hi + off);
  uint32_t l = *reinterpret_cast<const uint32_t*>(lo + off);
  float2 o;
  o.x = __uint_as_float(h << 16) + __uint_as_float(l << 16);
  o.y = __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u);
  return o;
}

// ----------------------------------------------------------------------------- forward
// shared memory (bytes, after 1024-alignment): W2 hi/lo, W3 hi/lo (4 x 8 KB) | tile hi/lo (2 x 16 KB) | floats
#define NB_EFT_W 0
#define NB_EFT_T (4 * NB_TC_TILE_BYTES(64))
#define NB_EFT_F (NB_EFT_T + 2 * NB_TC_TILE_BYTES(128))
#define NB_EFT_NFLOAT (4 * NB_H + NB_MAX_EF * NB_H + 2 * NB_TILE + 3 * NB_TILE)
#define NB_EDGE_FWD_TC_SMEM (NB_EFT_F + NB_EFT_NFLOAT * 4 + 64 + 1024)

__global__ void __launch_bounds__(NB_THREADS) k_edge_fwd_tc(NbEdgeFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char nb_smraw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)nb_smraw + 1023) & ~(uintptr_t)1023);
  unsigned char* W2h = base + NB_EFT_W;
  unsigned char* W2l = W2h + NB_TC_TILE_BYTES(64);
  unsigned char* W3h = W2l + NB_TC_TILE_BYTES(64);
  unsigned char* W3l = W3h + NB_TC_TILE_BYTES(64);
  unsigned char* Th = base + NB_EFT_T;
  unsigned char* Tl = Th + NB_TC_TILE_BYTES(128);
  float* fl = reinterpret_cast<float*>(base + NB_EFT_F);
  float* vb2 = fl;
  float* vb3 = vb2 + NB_H;
  float* vw4 = vb3 + NB_H;
  float* vwr = vw4 + NB_H;
  float* vwe = vwr + NB_H;                 // [NB_MAX_EF][64]
  float* cpart = vwe + NB_MAX_EF * NB_H;   // [2][128]
  float* rF = cpart + 2 * NB_TILE;         // [3][128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(rF + 3 * NB_TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const NbEdgeGeom g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;
  const int row = 32 * q + lane;
  const int cb = 32 * hf;

  nb_tc_stage_weight(W2h, W2l, a.w.W2, tid);
  nb_tc_stage_weight(W3h, W3l, a.w.W3, tid);
  if (tid < NB_H) {
    vb2[tid] = __ldg(a.w.b2 + tid);
    vb3[tid] = __ldg(a.w.b3 + tid);
    vw4[tid] = __ldg(a.w.w4 + tid);
    vwr[tid] = __ldg(a.w.W1 + (int64_t)tid * a.w.ldw1 + a.w.col_rad);
#pragma unroll
    for (int f = 0; f < NB_MAX_EF; ++f)
      vwe[f * NB_H + tid] = f < g.nef ? __ldg(a.w.W1 + (int64_t)tid * a.w.ldw1 + a.w.col_ef + f) : 0.f;
  }
  if (tid == 0) {
    nb_mbar_init(bar, 1);
    nb_mbar_fence_init();
  }
  if (warp == 0) nb_tmem_alloc(tmem_slot, 64);
  nb_fence_async_smem();
  nb_tc_fence_before();
  __syncthreads();
  nb_tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_mine = tm + ((uint32_t)(32 * q) << 16) + (uint32_t)cb;
  const uint32_t idesc = nb_idesc_bf16(128, 64, 0, 0);
  const float b4 = __ldg(a.w.b4);
  const int Nm1 = g.N - 1;
  uint32_t phase = 0;

  for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
    const int gt0 = u * g.G;
    const int ngt = min(g.G, g.NGT - gt0);
    const int R = ngt * g.EPG;
    const int64_t node0 = (int64_t)gt0 * g.N;
    for (int r0 = 0; r0 < R; r0 += NB_TILE) {
      const int nv = min(NB_TILE, R - r0);
      const NbRowRegs rr = nb_row_regs(g, a.x, a.ef, gt0, r0, nv, row);
      // z1 = SiLU(pre1) -> tile
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float v[8];
        nb_pre1_8(a.P, a.Q, rr, cb + 8 * jj, vwr, vwe, g.nef, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = nb_silu(v[i]);
        nb_tc_store8(Th, Tl, row, 4 * hf + jj, v);
      }
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        nb_tc_fence_after();
        nb_tc_issue3(tm, nb_smem_u32(Th), nb_smem_u32(Tl), 0, nb_smem_u32(W2h), nb_smem_u32(W2l), 0, 4, idesc, false);
        nb_mma_commit(bar);
      }
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      // m = SiLU(pre2 + b2) -> tile (z1 is no longer needed)
      {
        float v[32];
        nb_tmem_ld32(tm_mine, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = nb_silu(v[i] + vb2[cb + i]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) nb_tc_store8(Th, Tl, row, 4 * hf + jj, v + 8 * jj);
      }
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        nb_tc_fence_after();
        nb_tc_issue3(tm, nb_smem_u32(Th), nb_smem_u32(Tl), 0, nb_smem_u32(W3h), nb_smem_u32(W3l), 0, 4, idesc, false);
        nb_mma_commit(bar);
      }
      // while the tensor core runs phi_x: M_i += column sums of m over each receiver's rows (warp = receiver slot)
      const int s0 = r0 / Nm1, s1 = (r0 + nv - 1) / Nm1;
      for (int s = s0 + warp; s <= s1; s += 8) {
        int ra = max(s * Nm1, r0) - r0, rb = min((s + 1) * Nm1, r0 + nv) - r0;
        float2 sum = make_float2(0.f, 0.f);
        for (int r = ra; r < rb; ++r) {
          float2 t = nb_tc_load_pair(Th, Tl, r, lane);
          sum.x += t.x;
          sum.y += t.y;
        }
        float2* dst = reinterpret_cast<float2*>(a.M + (node0 + s) * NB_H + 2 * lane);
        if (s * Nm1 >= r0) *dst = sum;
        else {
          float2 old = *dst;
          *dst = make_float2(old.x + sum.x, old.y + sum.y);
        }
      }
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      // c = w4 . SiLU(pre3 + b3) + b4
      {
        float v[32];
        nb_tmem_ld32(tm_mine, v);
        float cp = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) cp = fmaf(vw4[cb + i], nb_silu(v[i] + vb3[cb + i]), cp);
        cpart[hf * NB_TILE + row] = cp;
      }
      nb_tc_fence_before();
      __syncthreads();
      if (hf == 0) {
        float c = cpart[row] + cpart[NB_TILE + row] + b4;
        float fx = rr.dx * c, fy = rr.dy * c, fz = rr.dz * c;
        if (g.clamp_edge) {
          fx = fminf(fmaxf(fx, -100.f), 100.f);
          fy = fminf(fmaxf(fy, -100.f), 100.f);
          fz = fminf(fmaxf(fz, -100.f), 100.f);
        }
        rF[row] = fx;
        rF[NB_TILE + row] = fy;
        rF[2 * NB_TILE + row] = fz;
      }
      __syncthreads();
      for (int idx = tid; idx < (s1 - s0 + 1) * 3; idx += NB_THREADS) {
        int s = s0 + idx / 3, d = idx % 3;
        int ra = max(s * Nm1, r0) - r0, rb = min((s + 1) * Nm1, r0 + nv) - r0;
        float sum = 0.f;
        for (int r = ra; r < rb; ++r) sum += rF[d * NB_TILE + r];
        float* dst = a.Fsum + (node0 + s) * 3 + d;
        if (s * Nm1 >= r0) *dst = sum;
        else *dst += sum;
      }
      __syncthreads();
    }
  }
  nb_tc_fence_before();
  __syncthreads();
  if (warp == 0) nb_tmem_dealloc(tm, 64);
}
#endif  // NB_EMU
