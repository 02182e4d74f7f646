// nb_segno_fused.cuh — SEGNO's T second-order integration sub-steps fused into ONE kernel (forward).
//
// Reference: SEGNO.forward_step (SEGNO/models/model.py:95-102) calls the weight-shared SEGNO_GCL T times
// (gcl.py:111-119).  Here a CTA owns a unit of G trajectories (G*N <= 27 nodes) for the whole call: the node state
// (h, x, v) lives in shared memory across all T sub-steps, every weight matrix is staged once as a split-bf16
// operand tile, and each sub-step runs on the tensor cores end to end:
//
//   P|Q      = h W1[:, h_row]^T + b1 | h W1[:, h_col]^T          node MMA (M = 64)        -> node tile
//   per 128-edge tile: gather -> SiLU -> W2 -> SiLU -> W3 -> phi_x head -> scatter           (as k_edge_fwd_sel)
//   a        = clamp-per-edge mean force ; v += a/T ; x += v/T                              (gcl.py:100-102,116-117)
//   U5       = [h, M] W5^T + b5 ; h <- h + SiLU(U5) W6^T + b6                               node MMAs (gcl.py:89-94)
//
// When `saved` is given, the per-sub-step state the (unfused) backward needs — h_k, M_k, U5_k, x_k — is written
// in the layout of nb_segno_forward's saved buffer; nothing else reaches HBM between sub-steps.
#pragma once
#ifndef NB_EMU
#include "nb_edge_sel.cuh"

struct NbSegnoFusedArgs {
  NbEdgeGeom g;       // NGT = B, clamp_edge = 1
  int T, recurrent;
  float inv_T, cw;
  const float *W1, *b1;   // edge_mlp.0 [64][ldw1]: cols h_row | h_col | radial | edge_attr
  int ldw1, col_rad, col_ef;
  const float *W2, *b2, *W3, *b3, *w4, *b4;   // edge_mlp.2, coord_mlp.0, coord_mlp.2
  const float *W5, *b5, *W6, *b6;             // node_mlp.0 [64][128], node_mlp.2 [64][64]
  const float *h_in, *x_in, *v_in;            // [B*N][64], [B*N][3], [B*N][3]
  const float* ef;                            // [B*EPG][nef]
  float *h_out, *x_out, *v_out;
  float* saved;                               // nullable; [T][iter_stride]: h | M | U5 | x  (each [B*N] rows)
  int64_t iter_stride;
};

// shared memory map (bytes after 1024-alignment)
#define NB_FS_W 0                                            // 14 weight pieces x 8 KB: W2 W3 W1r W1c W5a W5b W6 (hi, lo)
#define NB_FS_T (14 * NB_TC_TILE_BYTES(64))                  // activation tile hi/lo      2 x 16 KB
#define NB_FS_SEL (NB_FS_T + 2 * NB_TC_TILE_BYTES(128))      // selector                   16 KB
#define NB_FS_NT (NB_FS_SEL + NB_TC_TILE_BYTES(128))         // node tile hi/lo            2 x 8 KB
#define NB_FS_F (NB_FS_NT + 2 * NB_TC_TILE_BYTES(64))        // force tile hi/lo           2 x 2 KB
#define NB_FS_HN (NB_FS_F + 2 * NB_TILE * 16)                // h tile hi/lo (64 rows)     2 x 8 KB
#define NB_FS_UN (NB_FS_HN + 2 * NB_TC_TILE_BYTES(64))       // M / SiLU(U5) tile hi/lo    2 x 8 KB
#define NB_FS_FL (NB_FS_UN + 2 * NB_TC_TILE_BYTES(64))
#define NB_FS_NFLOAT (28 * NB_H + 2 * 32 * 3 + 6 * NB_H + 2 * NB_TILE)
#define NB_SEGNO_FUSED_SMEM(RU) (NB_FS_FL + NB_FS_NFLOAT * 4 + (RU) * 4 + 64 + 1024)
#define NB_FS_TMEM_COLS 512  // [0,64) pre | [64,128) M sums | [128,136) F sums | [192,256) P / U5 / dh | [256,320) Q

__device__ __forceinline__ void nb_fs_stage_w(unsigned char* hi, unsigned char* lo, const float* __restrict__ W, int ldw,
                                              int tid) {
  for (int idx = tid; idx < 64 * 8; idx += NB_THREADS) {
    int o = idx >> 3, j = idx & 7;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(W + (int64_t)o * ldw + 8 * j + i);
    nb_tc_store8(hi, lo, o, j, v);
  }
}

__global__ void __launch_bounds__(NB_THREADS, 1) k_segno_fused_fwd(NbSegnoFusedArgs a) {
  NB_PDL_ENTER();
  extern __shared__ __align__(1024) unsigned char nb_smraw[];
  unsigned char* base = nb_smraw + ((1024u - (nb_smem_u32(nb_smraw) & 1023u)) & 1023u);
  unsigned char* Wt = base + NB_FS_W;  // piece p (hi at 2p, lo at 2p+1): 0 W2, 1 W3, 2 W1r, 3 W1c, 4 W5a, 5 W5b, 6 W6
  unsigned char* Th = base + NB_FS_T;
  unsigned char* Tl = Th + NB_TC_TILE_BYTES(128);
  unsigned char* Sel = base + NB_FS_SEL;
  unsigned char* Nh = base + NB_FS_NT;
  unsigned char* Nl = Nh + NB_TC_TILE_BYTES(64);
  unsigned char* Fh = base + NB_FS_F;
  unsigned char* Fl = Fh + NB_TILE * 16;
  unsigned char* Hh = base + NB_FS_HN;
  unsigned char* Hl = Hh + NB_TC_TILE_BYTES(64);
  unsigned char* Uh = base + NB_FS_UN;
  unsigned char* Ul = Uh + NB_TC_TILE_BYTES(64);
  float* fl = reinterpret_cast<float*>(base + NB_FS_FL);
  float* hs = fl;                    // [28][64] node state h (fp32), G*N <= 27
  float* xs = hs + 28 * NB_H;        // [32][3]
  float* vs = xs + 32 * 3;           // [32][3]
  float* vb1 = vs + 32 * 3;
  float* vb2 = vb1 + NB_H;
  float* vb3 = vb2 + NB_H;
  float* vw4 = vb3 + NB_H;
  float* vb5 = vw4 + NB_H;
  float* vb6 = vb5 + NB_H;
  float* cpart = vb6 + NB_H;         // [2][128]
  uint32_t* rowinfo = reinterpret_cast<uint32_t*>(cpart + 2 * NB_TILE);
  const NbEdgeGeom g = a.g;
  const int RU = g.G * g.EPG, GN = g.G * g.N;
  uint64_t* bar = reinterpret_cast<uint64_t*>(rowinfo + RU + (RU & 1));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;
  const int row = 32 * q + lane;
  const int cb = 32 * hf;
  const int nl = 16 * q + lane;      // node owned in the M = 64 accumulators (lanes < 16 only)
  const bool nown = lane < 16;

#define NB_FS_WH(p) (Wt + (2 * (p)) * NB_TC_TILE_BYTES(64))
#define NB_FS_WL(p) (Wt + (2 * (p) + 1) * NB_TC_TILE_BYTES(64))
  nb_fs_stage_w(NB_FS_WH(0), NB_FS_WL(0), a.W2, NB_H, tid);
  nb_fs_stage_w(NB_FS_WH(1), NB_FS_WL(1), a.W3, NB_H, tid);
  nb_fs_stage_w(NB_FS_WH(2), NB_FS_WL(2), a.W1, a.ldw1, tid);
  nb_fs_stage_w(NB_FS_WH(3), NB_FS_WL(3), a.W1 + NB_H, a.ldw1, tid);
  nb_fs_stage_w(NB_FS_WH(4), NB_FS_WL(4), a.W5, 2 * NB_H, tid);
  nb_fs_stage_w(NB_FS_WH(5), NB_FS_WL(5), a.W5 + NB_H, 2 * NB_H, tid);
  nb_fs_stage_w(NB_FS_WH(6), NB_FS_WL(6), a.W6, NB_H, tid);
  // zero: node tile, h tile, U tile (rows beyond the unit's nodes must stay finite zeros)
  for (int idx = tid; idx < 2 * NB_TC_TILE_BYTES(64) / 16; idx += NB_THREADS) reinterpret_cast<uint4*>(Nh)[idx] = make_uint4(0u, 0u, 0u, 0u);
  for (int idx = tid; idx < 4 * NB_TC_TILE_BYTES(64) / 16; idx += NB_THREADS) reinterpret_cast<uint4*>(Hh)[idx] = make_uint4(0u, 0u, 0u, 0u);
  nb_sel_build_rowinfo(rowinfo, g, tid, NB_THREADS);
  if (tid < NB_H) {
    vb1[tid] = __ldg(a.b1 + tid);
    vb2[tid] = __ldg(a.b2 + tid);
    vb3[tid] = __ldg(a.b3 + tid);
    vw4[tid] = __ldg(a.w4 + tid);
    vb5[tid] = __ldg(a.b5 + tid);
    vb6[tid] = __ldg(a.b6 + tid);
  }
  if (tid == 0) {
    nb_mbar_init(bar, 1);
    nb_mbar_fence_init();
  }
  if (warp == 0) nb_tmem_alloc(tmem_slot, NB_FS_TMEM_COLS);
  __syncthreads();
  for (int idx = tid; idx < 10 * 8; idx += NB_THREADS) {  // weight rows of the node tile (w_rad, w_ef)
    int k = idx >> 3, j = idx & 7;
    int f = (k >> 1) - 1;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c = 8 * j + i;
      v[i] = f < 0 ? __ldg(a.W1 + (int64_t)c * a.ldw1 + a.col_rad)
                   : (f < g.nef ? __ldg(a.W1 + (int64_t)c * a.ldw1 + a.col_ef + f) : 0.f);
    }
    nb_tc_store8(Nh, Nl, NB_SEL_XC0 + k, j, v);
  }
  nb_fence_async_smem();
  nb_tc_fence_before();
  __syncthreads();
  nb_tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(32 * q) << 16;
  const uint32_t tm_mine = tm + lane_base + (uint32_t)cb;
  const uint32_t idesc_fwd = nb_idesc_bf16(128, 64, 0, 0);
  const uint32_t idesc_gat = nb_idesc_bf16(128, 64, 0, 1);
  const uint32_t idesc_sc = nb_idesc_bf16(64, 64, 1, 1);
  const uint32_t idesc_sc8 = nb_idesc_bf16(64, 8, 1, 1);
  const uint32_t idesc_node = nb_idesc_bf16(64, 64, 0, 0);   // node rows x W^T, M = 64
  const uint32_t sTh = nb_smem_u32(Th), sTl = nb_smem_u32(Tl), sSel = nb_smem_u32(Sel), sNh = nb_smem_u32(Nh),
                 sNl = nb_smem_u32(Nl), sFh = nb_smem_u32(Fh), sFl = nb_smem_u32(Fl), sHh = nb_smem_u32(Hh),
                 sHl = nb_smem_u32(Hl), sUh = nb_smem_u32(Uh), sUl = nb_smem_u32(Ul), sW = nb_smem_u32(Wt);
#define NB_FS_SWH(p) (sW + (2 * (p)) * NB_TC_TILE_BYTES(64))
#define NB_FS_SWL(p) (sW + (2 * (p) + 1) * NB_TC_TILE_BYTES(64))
  const float b4 = __ldg(a.b4);
  const float cnt = (float)(g.N - 1 > 1 ? g.N - 1 : 1);
  const int64_t Nn = (int64_t)g.NGT * g.N;
  uint32_t phase = 0;

  for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
    const int gt0 = u * g.G;
    const int ngt = min(g.G, g.NGT - gt0);
    const int R = ngt * g.EPG;
    const int nnode = ngt * g.N;
    const int64_t node0 = (int64_t)gt0 * g.N;
    // ---- load the unit's state
    for (int idx = tid; idx < nnode * 16; idx += NB_THREADS)
      nb_st4(hs + idx * 4, nb_ld4(a.h_in + node0 * NB_H + idx * 4));
    for (int idx = tid; idx < nnode * 3; idx += NB_THREADS) {
      xs[idx] = __ldg(a.x_in + node0 * 3 + idx);
      vs[idx] = __ldg(a.v_in + node0 * 3 + idx);
    }
    __syncthreads();

    for (int it = 0; it < a.T; ++it) {
      float* sv = a.saved ? a.saved + (int64_t)it * a.iter_stride : nullptr;  // h | M | U5 | x
      // ---- (a) h tile <- split(h) ; save h_k, x_k
      for (int idx = tid; idx < nnode * 8; idx += NB_THREADS) {
        int n = idx >> 3, j = idx & 7;
        float4 p0 = nb_ld4(hs + n * NB_H + 8 * j), p1 = nb_ld4(hs + n * NB_H + 8 * j + 4);
        float v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        nb_tc_store8(Hh, Hl, n, j, v);
        if (sv) {
          nb_st4(sv + (node0 + n) * NB_H + 8 * j, p0);
          nb_st4(sv + (node0 + n) * NB_H + 8 * j + 4, p1);
        }
      }
      if (sv)
        for (int idx = tid; idx < nnode * 3; idx += NB_THREADS) sv[3 * Nn * NB_H + node0 * 3 + idx] = xs[idx];
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      // ---- (b) P = h W1r^T, Q = h W1c^T -> node tile rows [0, GN) and [GN, 2 GN)
      if (NB_ISSUER(0)) {
        nb_tc_fence_after();
        nb_issue_w3(tm + 192, sHh, sHl, NB_FS_SWH(2), NB_FS_SWL(2), false, idesc_node, 0u);
        nb_issue_w3(tm + 256, sHh, sHl, NB_FS_SWH(3), NB_FS_SWL(3), false, idesc_node, 0u);
        nb_mma_commit(bar);
      }
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      {
        float v[32];
        nb_tmem_ld32(tm + lane_base + 192 + (uint32_t)cb, v);
        if (nown && nl < nnode) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += vb1[cb + i];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) nb_tc_store8(Nh, Nl, nl, 4 * hf + jj, v + 8 * jj);
        }
        nb_tmem_ld32(tm + lane_base + 256 + (uint32_t)cb, v);
        if (nown && nl < nnode) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) nb_tc_store8(Nh, Nl, GN + nl, 4 * hf + jj, v + 8 * jj);
        }
      }
      nb_tc_fence_before();
      // (the selector write below is followed by fence + __syncthreads before the gather reads the node tile)

      // ---- (c) edge tiles
      for (int r0 = 0; r0 < R; r0 += NB_TILE) {
        const int nv = min(NB_TILE, R - r0);
        const bool valid = row < nv;
        float dx = 0.f, dy = 0.f, dz = 0.f, r2 = 0.f;
        float e[NB_MAX_EF];
#pragma unroll
        for (int f = 0; f < NB_MAX_EF; ++f) e[f] = 0.f;
        int li = 0, lj = 0;
        if (valid) {
          const uint32_t ri = rowinfo[r0 + row];
          li = ri & 0xff;
          lj = (ri >> 8) & 0xff;
          const int lg = ri >> 16;
          dx = xs[li * 3 + 0] - xs[lj * 3 + 0];
          dy = xs[li * 3 + 1] - xs[lj * 3 + 1];
          dz = xs[li * 3 + 2] - xs[lj * 3 + 2];
          r2 = dx * dx + dy * dy + dz * dz;
          const int64_t eoff = ((int64_t)((gt0 + lg) % g.B) * g.EPG + (r0 + row - lg * g.EPG)) * g.nef;
#pragma unroll
          for (int f = 0; f < NB_MAX_EF; ++f)
            if (f < g.nef) e[f] = __ldg(a.ef + eoff + f);
        }
        if (r0 > 0) {  // the scatter MMAs of the previous tile have consumed Sel / T / F
          nb_mbar_wait(bar, phase);
          phase ^= 1;
          nb_tc_fence_after();
        }
        nb_sel_write_row(Sel, row, hf, valid, li, GN + lj, r2, e);
        nb_fence_async_smem();
        nb_tc_fence_before();
        __syncthreads();
        if (NB_ISSUER(0)) {
          nb_tc_fence_after();
          nb_issue_gather(tm, sSel, sNh, sNl, 4, idesc_gat, 0u);
          nb_mma_commit(bar);
        }
        nb_mbar_wait(bar, phase);
        phase ^= 1;
        nb_tc_fence_after();
        {
          float v[32];
          nb_tmem_ld32(tm_mine, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = nb_silu(v[i]);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) nb_tc_store8(Th, Tl, row, 4 * hf + jj, v + 8 * jj);
        }
        nb_fence_async_smem();
        nb_tc_fence_before();
        __syncthreads();
        if (NB_ISSUER(0)) {
          nb_tc_fence_after();
          nb_issue_w3(tm, sTh, sTl, NB_FS_SWH(0), NB_FS_SWL(0), false, idesc_fwd, 0u);
          nb_mma_commit(bar);
        }
        nb_mbar_wait(bar, phase);
        phase ^= 1;
        nb_tc_fence_after();
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
        if (NB_ISSUER(0)) {
          nb_tc_fence_after();
          nb_issue_w3(tm, sTh, sTl, NB_FS_SWH(1), NB_FS_SWL(1), false, idesc_fwd, 0u);
          nb_mma_commit(bar);
          nb_issue_scatter(tm + 64, sSel, sTh, sTl, idesc_sc, r0 > 0 ? 1u : 0u);  // M_i += Sel^T m
        }
        nb_mbar_wait(bar, phase);
        phase ^= 1;
        nb_tc_fence_after();
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
          float fx = dx * c, fy = dy * c, fz = dz * c;
          if (g.clamp_edge) {
            fx = fminf(fmaxf(fx, -100.f), 100.f);
            fy = fminf(fmaxf(fy, -100.f), 100.f);
            fz = fminf(fmaxf(fz, -100.f), 100.f);
          }
          const uint32_t px = nb_pack_split(fx), py = nb_pack_split(fy), pz = nb_pack_split(fz);
          *reinterpret_cast<uint4*>(Fh + row * 16) = make_uint4((px & 0xffffu) | (py << 16), pz & 0xffffu, 0u, 0u);
          *reinterpret_cast<uint4*>(Fl + row * 16) = make_uint4((px >> 16) | (py & 0xffff0000u), pz >> 16, 0u, 0u);
        }
        nb_fence_async_smem();
        __syncthreads();
        if (NB_ISSUER(0)) {
          nb_tc_fence_after();
          nb_issue_scatter8(tm + 128, sSel, sFh, sFl, idesc_sc8, r0 > 0 ? 1u : 0u);
          nb_mma_commit(bar);
        }
      }
      // ---- (d) unit read-out: M_i -> U tile (+ saved), Fsum_i -> integrator
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      {
        float v[32];
        nb_tmem_ld32(tm + lane_base + 64 + (uint32_t)cb, v);
        if (nown && nl < nnode) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) nb_tc_store8(Uh, Ul, nl, 4 * hf + jj, v + 8 * jj);
          if (sv) {
            float* dst = sv + Nn * NB_H + (node0 + nl) * NB_H + cb;
#pragma unroll
            for (int k = 0; k < 8; ++k) nb_st4(dst + 4 * k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
          }
        }
        float f4[4];
        nb_tmem_ld4(tm + lane_base + 128, f4);
        if (hf == 0 && nown && nl < nnode) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const float acc = f4[d] / cnt * a.cw;             // gcl.py:101-102
            const float vn = vs[nl * 3 + d] + acc * a.inv_T;  // gcl.py:116
            vs[nl * 3 + d] = vn;
            xs[nl * 3 + d] = xs[nl * 3 + d] + vn * a.inv_T;   // gcl.py:117
          }
        }
      }
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      // ---- (e) U5 = [h, M] W5^T + b5 ; h <- h + SiLU(U5) W6^T + b6
      if (NB_ISSUER(0)) {
        nb_tc_fence_after();
        nb_issue_w3(tm + 192, sHh, sHl, NB_FS_SWH(4), NB_FS_SWL(4), false, idesc_node, 0u);
        nb_issue_w3(tm + 192, sUh, sUl, NB_FS_SWH(5), NB_FS_SWL(5), false, idesc_node, 1u);
        nb_mma_commit(bar);
      }
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      {
        float v[32];
        nb_tmem_ld32(tm + lane_base + 192 + (uint32_t)cb, v);
        if (nown && nl < nnode) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += vb5[cb + i];
          if (sv) {
            float* dst = sv + 2 * Nn * NB_H + (node0 + nl) * NB_H + cb;
#pragma unroll
            for (int k = 0; k < 8; ++k) nb_st4(dst + 4 * k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = nb_silu(v[i]);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) nb_tc_store8(Uh, Ul, nl, 4 * hf + jj, v + 8 * jj);
        }
      }
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      if (NB_ISSUER(0)) {
        nb_tc_fence_after();
        nb_issue_w3(tm + 192, sUh, sUl, NB_FS_SWH(6), NB_FS_SWL(6), false, idesc_node, 0u);
        nb_mma_commit(bar);
      }
      nb_mbar_wait(bar, phase);
      phase ^= 1;
      nb_tc_fence_after();
      {
        float v[32];
        nb_tmem_ld32(tm + lane_base + 192 + (uint32_t)cb, v);
        if (nown && nl < nnode) {
          float* hp = hs + nl * NB_H + cb;
#pragma unroll
          for (int i = 0; i < 32; ++i) hp[i] = (a.recurrent ? hp[i] : 0.f) + (v[i] + vb6[cb + i]);
        }
      }
      nb_tc_fence_before();
      __syncthreads();
    }
    // ---- write the unit's final state
    for (int idx = tid; idx < nnode * 16; idx += NB_THREADS) nb_st4(a.h_out + node0 * NB_H + idx * 4, nb_ld4(hs + idx * 4));
    for (int idx = tid; idx < nnode * 3; idx += NB_THREADS) {
      a.x_out[node0 * 3 + idx] = xs[idx];
      a.v_out[node0 * 3 + idx] = vs[idx];
    }
    __syncthreads();
  }
  nb_tc_fence_before();
  __syncthreads();
  if (warp == 0) nb_tmem_dealloc(tm, NB_FS_TMEM_COLS);
}
#endif  // NB_EMU
