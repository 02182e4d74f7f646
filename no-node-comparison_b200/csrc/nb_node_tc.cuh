// nb_node_tc.cuh — node-level 64-wide GEMMs and weight-gradient reductions on tcgen05 (split-bf16 operands, fp32
// accumulation in TMEM).  Same argument structs and semantics as k_gemm64 / k_wgrad64 (nb_node.cuh), which remain
// the variants the host emulator runs.
#pragma once
#ifndef NB_EMU
#include "nb_node.cuh"
#include "nb_tc.cuh"
#include "nb_edge_sel.cuh"

// ============================================================================= gemm64 on tensor cores
// Persistent CTAs: blockIdx.y = job, blockIdx.x strides over the job's 128-row tiles.  The B operands (weights) are
// split and staged once per CTA; per tile the A rows are loaded (coalesced, 32 bytes per thread), split into their
// bf16 pieces and multiplied with three MMA passes per source into one TMEM accumulator.
// Shared memory: per source an A tile [128][64] and a B tile [64][64], hi + lo (48 KB per source).
#define NB_GT_A(s) ((s) * (2 * NB_TC_TILE_BYTES(128) + 2 * NB_TC_TILE_BYTES(64)))
#define NB_GT_B(s) (NB_GT_A(s) + 2 * NB_TC_TILE_BYTES(128))
#define NB_GEMM_TC_SMEM(nsrc) ((nsrc) * (2 * NB_TC_TILE_BYTES(128) + 2 * NB_TC_TILE_BYTES(64)) + 64 + 1024)

__global__ void __launch_bounds__(NB_THREADS, 3) k_gemm64_tc(NbGemmBatch batch, int nsrc_max) {
  const NbGemmArgs& a = batch.job[blockIdx.y];
  const int ntiles = (a.rows + NB_TILE - 1) / NB_TILE;
  if ((int)blockIdx.x >= ntiles) return;
  extern __shared__ __align__(1024) unsigned char nb_smraw[];
  unsigned char* base = nb_smraw + ((1024u - (nb_smem_u32(nb_smraw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + NB_GT_A(nsrc_max));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;
  const int row = 32 * q + lane, cb = 32 * hf;

  if (tid == 0) {
    nb_mbar_init(bar, 1);
    nb_mbar_fence_init();
  }
  if (warp == 0) nb_tmem_alloc(tmem_slot, 64);
  for (int s = 0; s < a.nsrc; ++s) {
    const NbGemmSrc src = a.src[s];
    unsigned char* Bh = base + NB_GT_B(s);
    unsigned char* Bl = Bh + NB_TC_TILE_BYTES(64);
    // B operand, K-major: tile row n (output column), tile column k:  W[k * sk + n * sn] * scale
    for (int idx = tid; idx < 64 * 8; idx += NB_THREADS) {
      int n, j;
      if (src.sn == 1) { n = idx & 63; j = idx >> 6; }   // adjacent threads -> adjacent n (contiguous when sn == 1)
      else { n = idx >> 3; j = idx & 7; }                // adjacent threads -> adjacent k chunks (contiguous when sk == 1)
      float v[8];
      const float* w = src.W + (int64_t)n * src.sn + (int64_t)(8 * j) * src.sk;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 8 * j + i < src.kmax ? __ldg(w + (int64_t)i * src.sk) * src.scale : 0.f;
      nb_tc_store8(Bh, Bl, n, j, v);
    }
  }
  nb_tc_fence_before();
  __syncthreads();
  nb_tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t idesc = nb_idesc_bf16(128, 64, 0, 0);
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * NB_TILE;
    const int nv = min(NB_TILE, a.rows - r0);
    for (int s = 0; s < a.nsrc; ++s) {
      const NbGemmSrc src = a.src[s];
      unsigned char* Ah = base + NB_GT_A(s);
      unsigned char* Al = Ah + NB_TC_TILE_BYTES(128);
      // A operand, K-major: 8 threads per row, 32 contiguous bytes each; all four loads of a thread are in flight together
      float4 x0[4], x1[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = tid + it * NB_THREADS;
        const int r = idx >> 3, j = idx & 7;
        if (r < nv) {
          const float* p = src.A + (int64_t)(r0 + r) * src.lda + 8 * j;
          x0[it] = nb_ld4(p);
          x1[it] = nb_ld4(p + 4);
        } else {
          x0[it] = x1[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = tid + it * NB_THREADS;
        const int r = idx >> 3, j = idx & 7;
        float v[8] = {x0[it].x, x0[it].y, x0[it].z, x0[it].w, x1[it].x, x1[it].y, x1[it].z, x1[it].w};
        if (src.a_silu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = nb_silu(v[i]);
        }
        nb_tc_store8(Ah, Al, r, j, v);
      }
    }
    nb_fence_async_smem();
    nb_tc_fence_before();
    __syncthreads();
    if (NB_ISSUER(0)) {
      nb_tc_fence_after();
      for (int s = 0; s < a.nsrc; ++s) {
        const uint32_t sa = nb_smem_u32(base + NB_GT_A(s)), sb = nb_smem_u32(base + NB_GT_B(s));
        nb_issue_w3(tm, sa, sa + NB_TC_TILE_BYTES(128), sb, sb + NB_TC_TILE_BYTES(64), false, idesc, s > 0 ? 1u : 0u);
      }
      nb_mma_commit(bar);
    }
    nb_mbar_wait(bar, phase);
    phase ^= 1;
    nb_tc_fence_after();
    {
      float v[32];
      nb_tmem_ld32(tm + ((uint32_t)(32 * q) << 16) + (uint32_t)cb, v);
      if (row < nv) {
        const int64_t gr = r0 + row;
        if (a.bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __ldg(a.bias + cb + i);  // parameter offsets are not 16-byte aligned
        }
        if (a.out_pre) {
          float* o = a.out_pre + gr * a.ldp + cb;
#pragma unroll
          for (int k = 0; k < 8; ++k) nb_st4(o + 4 * k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
        }
        if (a.epi == NB_EPI_SILU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = nb_silu(v[i]);
        } else if (a.epi == NB_EPI_MUL_DSILU) {
          const float* up = a.U + gr * a.ldu + cb;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float4 u = nb_ld4(up + 4 * k);
            v[4 * k + 0] *= nb_dsilu(u.x);
            v[4 * k + 1] *= nb_dsilu(u.y);
            v[4 * k + 2] *= nb_dsilu(u.z);
            v[4 * k + 3] *= nb_dsilu(u.w);
          }
        }
        if (a.R) {
          const float* rp = a.R + gr * a.ldr + cb;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float4 t = nb_ld4(rp + 4 * k);
            v[4 * k + 0] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
          }
        }
        if (a.out) {
          float* o = a.out + gr * a.ldo + cb;
          if (a.accumulate) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float4 t = nb_ld4(o + 4 * k);
              v[4 * k + 0] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) nb_st4(o + 4 * k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
        }
      }
    }
    nb_tc_fence_before();  // the next tile's MMA overwrites the accumulator this thread has just read
  }
  nb_tc_fence_before();
  __syncthreads();
  if (warp == 0) nb_tmem_dealloc(tm, 64);
}

// ============================================================================= wgrad64 on tensor cores
// partial[cta][o*64 + k] = sum over this CTA's rows of sum_p scale_p G_p[r][o] act_p(A_p[r][k]) ; [4096 + o] = colsum(G_0)
// D[64 x 64] = G^T A with both tiles MN-major (K = 128 rows), accumulated in TMEM over all tiles of the CTA.
#define NB_WT_SMEM (4 * NB_TC_TILE_BYTES(128) + NB_TILE * 16 + 64 + 1024)

__global__ void __launch_bounds__(NB_THREADS, 3) k_wgrad64_tc(NbWgradBatch batch) {
  const NbWgradArgs& a = batch.job[blockIdx.y];
  extern __shared__ __align__(1024) unsigned char nb_smraw[];
  unsigned char* base = nb_smraw + ((1024u - (nb_smem_u32(nb_smraw) & 1023u)) & 1023u);
  unsigned char* Gh = base;
  unsigned char* Gl = Gh + NB_TC_TILE_BYTES(128);
  unsigned char* Ah = Gl + NB_TC_TILE_BYTES(128);
  unsigned char* Al = Ah + NB_TC_TILE_BYTES(128);
  unsigned char* ones = Al + NB_TC_TILE_BYTES(128);
  uint64_t* bar = reinterpret_cast<uint64_t*>(ones + NB_TILE * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;

  if (tid < NB_TILE) {
    uint32_t one2 = 0x3F803F80u;
    *reinterpret_cast<uint4*>(ones + tid * 16) = make_uint4(one2, one2, one2, one2);
  }
  if (tid == 0) {
    nb_mbar_init(bar, 1);
    nb_mbar_fence_init();
  }
  if (warp == 0) nb_tmem_alloc(tmem_slot, 128);
  nb_tc_fence_before();
  __syncthreads();
  nb_tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t idesc_wg = nb_idesc_bf16(64, 64, 1, 1), idesc_bs = nb_idesc_bf16(64, 8, 1, 1);
  const uint32_t sGh = nb_smem_u32(Gh), sGl = nb_smem_u32(Gl), sAh = nb_smem_u32(Ah), sAl = nb_smem_u32(Al),
                 sOnes = nb_smem_u32(ones);
  uint32_t phase = 0, wacc = 0;
  const int ntiles = (a.rows + NB_TILE - 1) / NB_TILE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * NB_TILE;
    const int nv = min(NB_TILE, a.rows - r0);
    for (int p = 0; p < a.npair; ++p) {
      const NbWgradPair pr = a.pair[p];
      if (wacc) {  // the MMAs reading the tiles have completed
        nb_mbar_wait(bar, phase);
        phase ^= 1;
        nb_tc_fence_after();
      }
      for (int idx = tid; idx < NB_TILE * 8; idx += NB_THREADS) {
        const int r = idx >> 3, j = idx & 7;
        float gv[8], av[8];
        if (r < nv) {
          const float* gp = pr.G + (int64_t)(r0 + r) * pr.ldg + 8 * j;
          const float* ap = pr.A + (int64_t)(r0 + r) * pr.lda + 8 * j;
          float4 g0 = nb_ld4(gp), g1 = nb_ld4(gp + 4), a0 = nb_ld4(ap), a1 = nb_ld4(ap + 4);
          gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w; gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
          av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
#pragma unroll
          for (int i = 0; i < 8; ++i) gv[i] *= pr.scale;
          if (pr.a_silu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = nb_silu(av[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) gv[i] = av[i] = 0.f;
        }
        nb_tc_store8(Gh, Gl, r, j, gv);
        nb_tc_store8(Ah, Al, r, j, av);
      }
      nb_fence_async_smem();
      nb_tc_fence_before();
      __syncthreads();
      if (NB_ISSUER(0)) {
        nb_tc_fence_after();
        if (a.colsum && p == 0) {
          nb_issue_wgrad(tm, tm + 64, sGh, sGl, sAh, sAl, sOnes, idesc_wg, idesc_bs, wacc);
        } else {
          uint32_t acc = wacc;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t ga = nb_desc_lo_mn(pass == 1 ? sGl : sGh);
            const uint32_t ab = nb_desc_lo_mn(pass == 2 ? sAl : sAh);
#pragma unroll
            for (int s = 0; s < 8; ++s) {
              nb_mma2(tm, ga + NB_KSTEP_MN * s, NB_DESC_HI_SW128, ab + NB_KSTEP_MN * s, NB_DESC_HI_SW128, idesc_wg, acc);
              acc = 1u;
            }
          }
        }
        nb_mma_commit(bar);
      }
      wacc = 1;
    }
  }
  if (wacc) {
    nb_mbar_wait(bar, phase);
    phase ^= 1;
    nb_tc_fence_after();
  }
  float* out = a.partial + (int64_t)blockIdx.x * NB_WGRAD_PLEN;
  {
    // accumulator row o <-> TMEM lane (o % 16) + 32 (o / 16): thread (q, lane < 16) owns row 16 q + lane
    const int o = 16 * q + lane;
    float v[32];
    nb_tmem_ld32(tm + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * hf), v);
    if (lane < 16) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        nb_st4(out + o * NB_H + 32 * hf + 4 * k,
               wacc ? make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    float c4[4];
    nb_tmem_ld4(tm + ((uint32_t)(32 * q) << 16) + 64, c4);
    if (lane < 16 && hf == 0) out[NB_H * NB_H + o] = (a.colsum && wacc) ? c4[0] : 0.f;
  }
  nb_tc_fence_before();
  __syncthreads();
  if (warp == 0) nb_tmem_dealloc(tm, 128);
}
#endif  // NB_EMU
