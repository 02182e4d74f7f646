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
    int64_t eoff = ((int64_t)nb_ef_graph(g, gt) * g.EPG + rem) * g.nef;
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
  NB_PDL_ENTER();
  extern __shared__ __align__(1024) unsigned char nb_smraw[];
  // 1024-byte alignment by pointer arithmetic on the shared array (keeps the shared state space: LDS/STS, not LD/ST)
  unsigned char* base = nb_smraw + ((1024u - (nb_smem_u32(nb_smraw) & 1023u)) & 1023u);
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

// ----------------------------------------------------------------------------- backward
// Per 128-edge tile (see k_edge_bwd in nb_edge.cuh for the math):
//   recompute z1, m (MMA 1, 2)  ->  g3  ->  dgrad3 + wgrad3  ->  g2  ->  dgrad2 + wgrad2  ->  g1  -> node reductions.
// The data-gradient MMA of each stage is issued first and waited for; the weight-gradient MMA of the same stage
// (K = 128 rows, accumulating into TMEM across ALL tiles of the CTA) runs on the tensor pipe underneath the next
// stage's CUDA-core work.  SiLU' of layer 2 is parked in TMEM (tcgen05.st over the pre-activation it came from),
// SiLU' of layer 1 stays in registers; bias gradients are column sums taken by a tiny N = 8 MMA against a ones tile.
//
// TMEM columns: [0,64) pre2 -> SiLU'(pre2) | [64,128) pre3 -> gm -> gz1 | [128,192) dW3 | [192,256) dW2 |
//               [256,264) db3 | [264,272) db2          (M = 64 accumulators occupy lanes (i%16) + 32 (i/16))
#define NB_EBT_W 0
#define NB_EBT_TZ (4 * NB_TC_TILE_BYTES(64))
#define NB_EBT_TM (NB_EBT_TZ + 2 * NB_TC_TILE_BYTES(128))
#define NB_EBT_TG (NB_EBT_TM + 2 * NB_TC_TILE_BYTES(128))
#define NB_EBT_ONES (NB_EBT_TG + 2 * NB_TC_TILE_BYTES(128))
#define NB_EBT_F (NB_EBT_ONES + NB_TILE * 16)
#define NB_EBT_NFLOAT(GN) (4 * NB_H + NB_MAX_EF * NB_H + 2 * NB_TILE + 3 * NB_TILE + NB_TILE + NB_MAX_EF * NB_TILE + (GN) * (NB_H + 3))
#define NB_EDGE_BWD_TC_SMEM(GN) (NB_EBT_F + NB_EBT_NFLOAT(GN) * 4 + 64 + 1024)
#define NB_EBT_TMEM_COLS 512

// fp32 [128][64] scratch with 16-byte chunks XOR-swizzled by (row & 7): conflict-free for row-owner writes and
// for column-owner reads
__device__ __forceinline__ float* nb_scratch_chunk(float* S, int r, int j4) { return S + r * NB_H + ((j4 ^ (r & 7)) << 2); }
__device__ __forceinline__ float nb_scratch_get(const float* S, int r, int c) {
  return S[r * NB_H + ((((c >> 2) ^ (r & 7))) << 2) + (c & 3)];
}

__global__ void __launch_bounds__(NB_THREADS, 1) k_edge_bwd_tc(NbEdgeBwdArgs a) {
  NB_PDL_ENTER();
  extern __shared__ __align__(1024) unsigned char nb_smraw[];
  // 1024-byte alignment by pointer arithmetic on the shared array (keeps the shared state space: LDS/STS, not LD/ST)
  unsigned char* base = nb_smraw + ((1024u - (nb_smem_u32(nb_smraw) & 1023u)) & 1023u);
  unsigned char* W2h = base + NB_EBT_W;
  unsigned char* W2l = W2h + NB_TC_TILE_BYTES(64);
  unsigned char* W3h = W2l + NB_TC_TILE_BYTES(64);
  unsigned char* W3l = W3h + NB_TC_TILE_BYTES(64);
  unsigned char* Tzh = base + NB_EBT_TZ;
  unsigned char* Tzl = Tzh + NB_TC_TILE_BYTES(128);
  unsigned char* Tmh = base + NB_EBT_TM;
  unsigned char* Tml = Tmh + NB_TC_TILE_BYTES(128);
  unsigned char* Tgh = base + NB_EBT_TG;
  unsigned char* Tgl = Tgh + NB_TC_TILE_BYTES(128);
  unsigned char* ones = base + NB_EBT_ONES;
  float* scratch = reinterpret_cast<float*>(Tmh);  // fp32 [128][64] view of the m tile area (free once dW3 is done)
  float* fl = reinterpret_cast<float*>(base + NB_EBT_F);
  float* vb2 = fl;
  float* vb3 = vb2 + NB_H;
  float* vw4 = vb3 + NB_H;
  float* vwr = vw4 + NB_H;
  float* vwe = vwr + NB_H;                    // [NB_MAX_EF][64]
  float* cpart = vwe + NB_MAX_EF * NB_H;      // [2][128]
  float* rG = cpart + 2 * NB_TILE;            // [3][128] dL/drij
  float* rR2 = rG + 3 * NB_TILE;              // [128]
  float* rE = rR2 + NB_TILE;                  // [NB_MAX_EF][128]
  float* gQacc = rE + NB_MAX_EF * NB_TILE;    // [G*N][64]
  float* gxacc = gQacc + a.g.G * a.g.N * NB_H;  // [G*N][3]
  uint64_t* bar = reinterpret_cast<uint64_t*>(gxacc + a.g.G * a.g.N * 3 + ((a.g.G * a.g.N * 3) & 1));
  uint64_t* bar2 = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar2 + 1);

  const NbEdgeGeom g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;
  const int row = 32 * q + lane;
  const int cb = 32 * hf;

  nb_tc_stage_weight(W2h, W2l, a.w.W2, tid);
  nb_tc_stage_weight(W3h, W3l, a.w.W3, tid);
  if (tid < NB_TILE) {
    uint32_t one2 = 0x3F803F80u;  // bf16 (1.0, 1.0)
    *reinterpret_cast<uint4*>(ones + tid * 16) = make_uint4(one2, one2, one2, one2);
  }
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
    nb_mbar_init(bar2, 1);
    nb_mbar_fence_init();
  }
  if (warp == 0) nb_tmem_alloc(tmem_slot, NB_EBT_TMEM_COLS);
  nb_fence_async_smem();
  nb_tc_fence_before();
  __syncthreads();
  nb_tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(32 * q) << 16;
  const uint32_t tA = tm + lane_base + 0 + (uint32_t)cb;    // pre2 / SiLU'(pre2)
  const uint32_t tB = tm + lane_base + 64 + (uint32_t)cb;   // pre3 / gm / gz1
  const uint32_t idesc_fwd = nb_idesc_bf16(128, 64, 0, 0);  // A K-major, B K-major
  const uint32_t idesc_dg = nb_idesc_bf16(128, 64, 0, 1);   // A K-major, B = W MN-major
  const uint32_t idesc_wg = nb_idesc_bf16(64, 64, 1, 1);    // A = g^T, B = act, both MN-major
  const uint32_t idesc_bs = nb_idesc_bf16(64, 8, 1, 1);     // column sums against the ones tile
  const float b4 = __ldg(a.w.b4);
  const int Nm1 = g.N - 1;
  uint32_t phase = 0, phase2 = 0;
  uint32_t wacc = 0;  // 0 until the weight-gradient accumulators hold their first tile

  float gw4acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) gw4acc[i] = 0.f;
  float gb4acc = 0.f;
  float gwr_acc = 0.f, gwe_acc[NB_MAX_EF];
#pragma unroll
  for (int f = 0; f < NB_MAX_EF; ++f) gwe_acc[f] = 0.f;
  const int rc = tid & 63, rpart = tid >> 6;  // reducer mapping: column, part (0..3)

  for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
    const int gt0 = u * g.G;
    const int ngt = min(g.G, g.NGT - gt0);
    const int R = ngt * g.EPG;
    const int nnode = ngt * g.N;
    const int64_t node0 = (int64_t)gt0 * g.N;
    for (int idx = tid; idx < nnode * NB_H; idx += NB_THREADS) gQacc[idx] = 0.f;
    for (int idx = tid; idx < nnode * 3; idx += NB_THREADS) gxacc[idx] = 0.f;

    for (int r0 = 0; r0 < R; r0 += NB_TILE) {
      const int nv = min(NB_TILE, R - r0);
      const NbRowRegs rr = nb_row_regs(g, a.x, a.ef, gt0, r0, nv, row);
      float gfx = 0.f, gfy = 0.f, gfz = 0.f;
      if (rr.valid) {
        gfx = __ldg(a.gFsum + (int64_t)rr.ni * 3 + 0);
        gfy = __ldg(a.gFsum + (int64_t)rr.ni * 3 + 1);
        gfz = __ldg(a.gFsum + (int64_t)rr.ni * 3 + 2);
      }
      if (hf == 0) {
        rR2[row] = rr.r2;
#pragma unroll
        for (int f = 0; f < NB_MAX_EF; ++f) rE[f * NB_TILE + row] = rr.e[f];
      }
      // ---- stage 1: z1 (+ SiLU' in registers) -> MMA 1
      float d1[32];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float v[8];
        nb_pre1_8(a.P, a.Q, rr, cb + 8 * jj, vwr, vwe, g.nef, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) nb_silu_grad(v[i], v[i], d1[8 * jj + i]);
        nb_tc_store8(Tzh, Tzl, row, 4 * hf + jj, v);
      }
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        nb_tc_fence_after();
        nb_tc_issue3(tm, nb_smem_u32(Tzh), nb_smem_u32(Tzl), 0, nb_smem_u32(W2h), nb_smem_u32(W2l), 0, 4, idesc_fwd, false);
        nb_mma_commit(bar);
      }
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      // ---- stage 2: m -> tile, SiLU'(pre2) -> TMEM, MMA 2
      {
        float v[32], d2[32];
        nb_tmem_ld32(tA, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) nb_silu_grad(v[i] + vb2[cb + i], v[i], d2[i]);
        nb_tmem_st32(tA, d2);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) nb_tc_store8(Tmh, Tml, row, 4 * hf + jj, v + 8 * jj);
      }
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        nb_tc_fence_after();
        nb_tc_issue3(tm + 64, nb_smem_u32(Tmh), nb_smem_u32(Tml), 0, nb_smem_u32(W3h), nb_smem_u32(W3l), 0, 4, idesc_fwd,
                     false);
        nb_mma_commit(bar);
      }
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      // ---- stage 3: phi_x head, g3 = dL/dpre3
      float cval;
      {
        float v[32], d3[32];
        nb_tmem_ld32(tB, v);
        float cp = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          nb_silu_grad(v[i] + vb3[cb + i], v[i], d3[i]);  // v = z3
          cp = fmaf(vw4[cb + i], v[i], cp);
        }
        cpart[hf * NB_TILE + row] = cp;
        __syncthreads();
        cval = cpart[row] + cpart[NB_TILE + row] + b4;
        if (g.clamp_edge) {  // clamp(rij * c) passes gradient only inside [-100, 100]
          float fx = rr.dx * cval, fy = rr.dy * cval, fz = rr.dz * cval;
          if (!(fx >= -100.f && fx <= 100.f)) gfx = 0.f;
          if (!(fy >= -100.f && fy <= 100.f)) gfy = 0.f;
          if (!(fz >= -100.f && fz <= 100.f)) gfz = 0.f;
        }
        const float gc = rr.dx * gfx + rr.dy * gfy + rr.dz * gfz;  // 0 for padded rows
        if (hf == 0) {
          rG[row] = cval * gfx;
          rG[NB_TILE + row] = cval * gfy;
          rG[2 * NB_TILE + row] = cval * gfz;
          gb4acc += gc;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          gw4acc[i] = fmaf(gc, v[i], gw4acc[i]);
          v[i] = gc * vw4[cb + i] * d3[i];  // g3
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) nb_tc_store8(Tgh, Tgl, row, 4 * hf + jj, v + 8 * jj);
      }
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        nb_tc_fence_after();
        // gm = g3 W3   (data gradient first: it is what the next stage waits for)
        nb_tc_issue3(tm + 64, nb_smem_u32(Tgh), nb_smem_u32(Tgl), 0, nb_smem_u32(W3h), nb_smem_u32(W3l), 1, 4, idesc_dg,
                     false);
        nb_mma_commit(bar);
        // dW3 += g3^T m ; db3 += g3^T 1
        nb_tc_issue3(tm + 128, nb_smem_u32(Tgh), nb_smem_u32(Tgl), 1, nb_smem_u32(Tmh), nb_smem_u32(Tml), 1, 8, idesc_wg,
                     wacc != 0);
        {
          uint32_t acc = wacc;
          for (int pass = 0; pass < 2; ++pass)
            for (int s = 0; s < 8; ++s) {
              nb_mma_bf16(tm + 256, nb_desc_mnmajor(nb_smem_u32(pass ? Tgl : Tgh), s),
                          nb_desc_mn8_noswizzle(nb_smem_u32(ones), s), idesc_bs, acc);
              acc = 1;
            }
        }
        nb_mma_commit(bar2);
      }
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      // ---- stage 4: g2 = (gm + gM_i) * SiLU'(pre2)
      {
        float v[32], d2[32];
        nb_tmem_ld32(tB, v);
        nb_tmem_ld32(tA, d2);
        if (rr.valid) {
          const float* gm = a.gM + (int64_t)rr.ni * NB_H + cb;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float4 t = nb_ld4(gm + 4 * k);
            v[4 * k + 0] = (v[4 * k + 0] + t.x) * d2[4 * k + 0];
            v[4 * k + 1] = (v[4 * k + 1] + t.y) * d2[4 * k + 1];
            v[4 * k + 2] = (v[4 * k + 2] + t.z) * d2[4 * k + 2];
            v[4 * k + 3] = (v[4 * k + 3] + t.w) * d2[4 * k + 3];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        nb_mbar_wait(bar2, phase2);  // dW3 / db3 have consumed the g3 and m tiles
        phase2 ^= 1;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) nb_tc_store8(Tgh, Tgl, row, 4 * hf + jj, v + 8 * jj);
      }
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        nb_tc_fence_after();
        nb_tc_issue3(tm + 64, nb_smem_u32(Tgh), nb_smem_u32(Tgl), 0, nb_smem_u32(W2h), nb_smem_u32(W2l), 1, 4, idesc_dg,
                     false);
        nb_mma_commit(bar);
        nb_tc_issue3(tm + 192, nb_smem_u32(Tgh), nb_smem_u32(Tgl), 1, nb_smem_u32(Tzh), nb_smem_u32(Tzl), 1, 8, idesc_wg,
                     wacc != 0);
        {
          uint32_t acc = wacc;
          for (int pass = 0; pass < 2; ++pass)
            for (int s = 0; s < 8; ++s) {
              nb_mma_bf16(tm + 264, nb_desc_mnmajor(nb_smem_u32(pass ? Tgl : Tgh), s),
                          nb_desc_mn8_noswizzle(nb_smem_u32(ones), s), idesc_bs, acc);
              acc = 1;
            }
        }
        nb_mma_commit(bar2);
      }
      wacc = 1;
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      // ---- stage 5: g1 = gz1 * SiLU'(pre1) -> fp32 scratch (the m tile area); dL/dr2 = w_rad . g1
      {
        float v[32];
        nb_tmem_ld32(tB, v);
        float gr2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] *= d1[i];
          gr2 = fmaf(vwr[cb + i], v[i], gr2);
        }
        cpart[hf * NB_TILE + row] = gr2;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          nb_st4(nb_scratch_chunk(scratch, row, (cb >> 2) + k), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
      }
      nb_tc_fence_before();
      __syncthreads();
      if (hf == 0) {
        float gr2 = cpart[row] + cpart[NB_TILE + row];
        rG[row] = fmaf(2.f * rr.dx, gr2, rG[row]);
        rG[NB_TILE + row] = fmaf(2.f * rr.dy, gr2, rG[NB_TILE + row]);
        rG[2 * NB_TILE + row] = fmaf(2.f * rr.dz, gr2, rG[2 * NB_TILE + row]);
      }
      __syncthreads();
      // ---- stage 6: node-level reductions (receiver rows -> gP and the w_rad / w_ef gradients, sender rows -> gQ, gx)
      const int s0 = r0 / Nm1, s1 = (r0 + nv - 1) / Nm1;
      {
        for (int s = s0 + rpart; s <= s1; s += 4) {
          int ra = max(s * Nm1, r0) - r0, rb = min((s + 1) * Nm1, r0 + nv) - r0;
          float sum = 0.f;
          for (int r = ra; r < rb; ++r) {
            float gv = nb_scratch_get(scratch, r, rc);
            sum += gv;
            gwr_acc = fmaf(gv, rR2[r], gwr_acc);
#pragma unroll
            for (int f = 0; f < NB_MAX_EF; ++f)
              if (f < g.nef) gwe_acc[f] = fmaf(gv, rE[f * NB_TILE + r], gwe_acc[f]);
          }
          float* dst = a.gP + (node0 + s) * NB_H + rc;
          if (s * Nm1 >= r0) *dst = sum;
          else *dst += sum;
        }
        const int v0 = (s0 / g.N) * g.N, v1 = (s1 / g.N + 1) * g.N;
        for (int v = v0 + rpart; v < v1; v += 4) {
          int lg = v / g.N, j = v - lg * g.N;
          int sa = max(s0, lg * g.N), sb = min(s1, lg * g.N + g.N - 1);
          float sum = 0.f;
          for (int s = sa; s <= sb; ++s) {
            int i = s - lg * g.N;
            if (i == j) continue;
            int r = s * Nm1 + (j < i ? j : j - 1) - r0;
            if (r < 0 || r >= nv) continue;
            sum += nb_scratch_get(scratch, r, rc);
          }
          gQacc[v * NB_H + rc] += sum;
        }
        for (int idx = tid; idx < (v1 - v0) * 3; idx += NB_THREADS) {
          int v = v0 + idx / 3, d = idx % 3;
          int lg = v / g.N, j = v - lg * g.N;
          float sum = 0.f;
          if (v >= s0 && v <= s1) {
            int ra = max(v * Nm1, r0) - r0, rb = min((v + 1) * Nm1, r0 + nv) - r0;
            for (int r = ra; r < rb; ++r) sum += rG[d * NB_TILE + r];
          }
          int sa = max(s0, lg * g.N), sb = min(s1, lg * g.N + g.N - 1);
          for (int s = sa; s <= sb; ++s) {
            int i = s - lg * g.N;
            if (i == j) continue;
            int r = s * Nm1 + (j < i ? j : j - 1) - r0;
            if (r < 0 || r >= nv) continue;
            sum -= rG[d * NB_TILE + r];
          }
          gxacc[v * 3 + d] += sum;
        }
      }
      nb_mbar_wait(bar2, phase2);  // dW2 / db2 have consumed the g2 and z1 tiles
      phase2 ^= 1;
      __syncthreads();
    }
    for (int idx = tid; idx < nnode * NB_H; idx += NB_THREADS) a.gQ[node0 * NB_H + idx] = gQacc[idx];
    for (int idx = tid; idx < nnode * 3; idx += NB_THREADS) a.gx[node0 * 3 + idx] += gxacc[idx];
    __syncthreads();
  }

  // ---- CTA epilogue: weight-gradient accumulators (TMEM) and register accumulators -> this CTA's partial slice
  float* out = a.partial + (int64_t)blockIdx.x * NB_EB_PLEN;
  nb_tc_fence_after();
  {
    // M = 64 accumulators: output row o lives in lane (o % 16) + 32 (o / 16)  ->  warp quarter q, lanes 0..15
    float v[32];
    const int o = 16 * q + lane;
    nb_tmem_ld32(tm + lane_base + 128 + (uint32_t)cb, v);  // dW3[o][cb..]
    if (lane < 16 && wacc) {
#pragma unroll
      for (int k = 0; k < 8; ++k) nb_st4(out + NB_EB_GW3 + o * NB_H + cb + 4 * k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
    }
    nb_tmem_ld32(tm + lane_base + 192 + (uint32_t)cb, v);  // dW2[o][cb..]
    if (lane < 16 && wacc) {
#pragma unroll
      for (int k = 0; k < 8; ++k) nb_st4(out + NB_EB_GW2 + o * NB_H + cb + 4 * k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
    }
    nb_tmem_ld32(tm + lane_base + 256, v);  // db3 in column 0, db2 in column 8
    if (lane < 16 && hf == 0 && wacc) {
      out[NB_EB_GB3 + o] = v[0];
      out[NB_EB_GB2 + o] = v[8];
    }
  }
  if (!wacc) {  // a CTA that processed no tile contributes zeros
    for (int idx = tid; idx < 2 * NB_H * NB_H + 2 * NB_H; idx += NB_THREADS) out[idx] = 0.f;
  }
  // dw4: sum of the per-row accumulators over the 128 rows (through the fp32 scratch); db4 likewise
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k)
    nb_st4(nb_scratch_chunk(scratch, row, (cb >> 2) + k),
           make_float4(gw4acc[4 * k], gw4acc[4 * k + 1], gw4acc[4 * k + 2], gw4acc[4 * k + 3]));
  if (hf == 0) cpart[row] = gb4acc;
  __syncthreads();
  {
    float s = 0.f;
    for (int r = rpart * 32; r < rpart * 32 + 32; ++r) s += nb_scratch_get(scratch, r, rc);
    float* red = rE;  // [4][64] (rE holds NB_MAX_EF * 128 >= 256 floats)
    red[rpart * NB_H + rc] = s;
    __syncthreads();
    if (tid < NB_H) out[NB_EB_GW4 + tid] = red[tid] + red[NB_H + tid] + red[2 * NB_H + tid] + red[3 * NB_H + tid];
    __syncthreads();
    red[rpart * NB_H + rc] = gwr_acc;
    __syncthreads();
    if (tid < NB_H) out[NB_EB_GWR + tid] = red[tid] + red[NB_H + tid] + red[2 * NB_H + tid] + red[3 * NB_H + tid];
#pragma unroll
    for (int f = 0; f < NB_MAX_EF; ++f) {
      __syncthreads();
      red[rpart * NB_H + rc] = gwe_acc[f];
      __syncthreads();
      if (tid < NB_H)
        out[NB_EB_GWE + f * NB_H + tid] = red[tid] + red[NB_H + tid] + red[2 * NB_H + tid] + red[3 * NB_H + tid];
    }
    if (tid < NB_H) {
      float t = 0.f;
      if (tid == 0)
        for (int r = 0; r < NB_TILE; ++r) t += cpart[r];
      out[NB_EB_GB4 + tid] = t;
    }
  }
  nb_tc_fence_before();
  __syncthreads();
  if (warp == 0) nb_tmem_dealloc(tm, NB_EBT_TMEM_COLS);
}
#endif  // NB_EMU
