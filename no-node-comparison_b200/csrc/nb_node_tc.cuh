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
#define NB_GEMM_TC_SMEM(nsrc) ((nsrc) * (2 * NB_TC_TILE_BYTES(128) + 2 * NB_TC_TILE_BYTES(64)) + 64 + NB_H * 4 + 1024)

// The A rows of tile n + 1 (and the epilogue's own global operands of tile n: U, R, the accumulate target) are requested
// BEFORE the CTA waits for the MMAs of tile n, so one HBM round trip per tile hides under the MMA + epilogue of the
// previous one instead of sitting in front of every tile (the kernel is HBM-latency bound: 2 x 32 B per thread in flight).
template <int NSRC>
__device__ __forceinline__ void nb_gemm_load_rows(const NbGemmArgs& a, int r0, int nv, int tid, float4 (&x0)[NSRC][4], float4 (&x1)[NSRC][4]) {
#pragma unroll
  for (int s = 0; s < NSRC; ++s) {
    if (s < a.nsrc) {
      const NbGemmSrc& src = a.src[s];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = tid + it * NB_THREADS;
        const int r = idx >> 3, j = idx & 7;
        if (r < nv) {
          const float* p = nb_seg_row(src.A, src.lda, src.seg_rows, src.seg_a, r0 + r) + 8 * j;
          x0[s][it] = nb_ld4(p);
          x1[s][it] = nb_ld4(p + 4);
        } else {
          x0[s][it] = x1[s][it] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
}

#ifdef NB_STAGE_CLOCKS
__device__ long long nb_dbg_clk_gemm[16];   // [0] launches sampled, [1..] cycles per stage of CTA (0, 0), thread 0
#define NB_GCLK(i)                                  \
  if (gdbg) {                                       \
    const long long t_now = clock64();              \
    atomicAdd((unsigned long long*)&nb_dbg_clk_gemm[i], (unsigned long long)(t_now - glast)); \
    glast = t_now;                                  \
  }
#else
#define NB_GCLK(i)
#endif

template <int NSRC>
__device__ __forceinline__ void nb_gemm64_tc_body(const NbGemmArgs& a, unsigned char* base, uint64_t* bar, uint32_t* tmem_slot, int nsrc_max) {
  const int ntiles = (a.rows + NB_TILE - 1) / NB_TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;
  const int row = 32 * q + lane, cb = 32 * hf;

#ifdef NB_STAGE_CLOCKS
  const bool gdbg = blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  long long glast = clock64();
  if (gdbg) atomicAdd((unsigned long long*)&nb_dbg_clk_gemm[0], 1ull);
#endif
  // the first tile's rows are in flight while the weights are staged
  float4 x0[NSRC][4], x1[NSRC][4];
  {
    const int r0 = blockIdx.x * NB_TILE;
    nb_gemm_load_rows<NSRC>(a, r0, min(NB_TILE, a.rows - r0), tid, x0, x1);
  }
  NB_GCLK(1)
  float* sbias = reinterpret_cast<float*>(tmem_slot + 2);   // [64]
  if (tid < NB_H) sbias[tid] = a.bias ? __ldg(a.bias + tid) : 0.f;
  if (tid == 0) {
    nb_mbar_init(bar, 1);
    nb_mbar_fence_init();
  }
  if (warp == 0) nb_tmem_alloc(tmem_slot, 64);
  NB_GCLK(2)
  for (int s = 0; s < a.nsrc; ++s) {
    const NbGemmSrc src = a.src[s];
    unsigned char* Bh = base + NB_GT_B(s);
    unsigned char* Bl = Bh + NB_TC_TILE_BYTES(64);
    if (src.img) {  // pre-split image of the weight block (hi tile, lo tile): a straight 16 KB copy
      const uint4* im = reinterpret_cast<const uint4*>(src.img);
      for (int idx = tid; idx < 2 * NB_TC_TILE_BYTES(64) / 16; idx += NB_THREADS) reinterpret_cast<uint4*>(Bh)[idx] = __ldg(im + idx);
      continue;
    }
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
  const uint32_t idesc_k = nb_idesc_bf16(128, 64, 0, 0), idesc_mn = nb_idesc_bf16(128, 64, 0, 1);
  uint32_t phase = 0;
  NB_GCLK(3)

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * NB_TILE;
    const int nv = min(NB_TILE, a.rows - r0);
#pragma unroll
    for (int s = 0; s < NSRC; ++s) {
      if (s < a.nsrc) {
        unsigned char* Ah = base + NB_GT_A(s);
        unsigned char* Al = Ah + NB_TC_TILE_BYTES(128);
        const bool a_silu = a.src[s].a_silu != 0;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int idx = tid + it * NB_THREADS;
          const int r = idx >> 3, j = idx & 7;
          float v[8] = {x0[s][it].x, x0[s][it].y, x0[s][it].z, x0[s][it].w, x1[s][it].x, x1[s][it].y, x1[s][it].z, x1[s][it].w};
          if (a_silu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = nb_silu(v[i]);
          }
          nb_tc_store8(Ah, Al, r, j, v);
        }
      }
    }
    NB_GCLK(4)
    nb_fence_async_smem();
    nb_tc_fence_before();
    __syncthreads();
    NB_GCLK(5)
    if (NB_ISSUER(0)) {
      nb_tc_fence_after();
      for (int s = 0; s < a.nsrc; ++s) {
        const uint32_t sa = nb_smem_u32(base + NB_GT_A(s)), sb = nb_smem_u32(base + NB_GT_B(s));
        const bool mn = a.src[s].img && a.src[s].img_mn;
        nb_issue_w3(tm, sa, sa + NB_TC_TILE_BYTES(128), sb, sb + NB_TC_TILE_BYTES(64), mn, mn ? idesc_mn : idesc_k, s > 0 ? 1u : 0u);
      }
      nb_mma_commit(bar);
    }
    NB_GCLK(6)
    // ---- while the MMAs run: the next tile's rows and this tile's epilogue operands
    {
      const int nt = tile + gridDim.x;
      if (nt < ntiles) nb_gemm_load_rows<NSRC>(a, nt * NB_TILE, min(NB_TILE, a.rows - nt * NB_TILE), tid, x0, x1);
    }
    const int64_t gr = r0 + row;
    const bool live = row < nv;
    // U (SiLU' epilogue), else R (residual), else the accumulate target: at most one of them is prefetched, and only by
    // single-source jobs (two-source jobs already hold 64 registers of next-tile rows)
    float4 eu[NSRC == 1 ? 8 : 1];
    const float* ep = nullptr;
    if (NSRC == 1 && live) {
      if (a.epi == NB_EPI_MUL_DSILU) ep = a.U + gr * a.ldu + cb;
      else if (a.R) ep = a.R + gr * a.ldr + cb;
      else if (a.out && a.accumulate) ep = a.out + gr * a.ldo + cb;
    }
    if (NSRC == 1 && ep) {
#pragma unroll
      for (int k = 0; k < (NSRC == 1 ? 8 : 1); ++k) eu[k] = nb_ld4(ep + 4 * k);
    }
    NB_GCLK(7)
    nb_mbar_wait(bar, phase);
    phase ^= 1;
    nb_tc_fence_after();
    NB_GCLK(8)
    {
      // Results leave through shared memory: a thread owns one row (128 B of it), so a direct store instruction of a warp
      // would touch 32 different lines with 16 B each; staged in the (now free) A tile of source 0 as [128][64] fp32 with
      // the 16-byte chunks XOR-swizzled by row, every store instruction of a warp then writes two whole rows (4 lines).
      float* stage = reinterpret_cast<float*>(base + NB_GT_A(0));
      float v[32];
      nb_tmem_ld32(tm + ((uint32_t)(32 * q) << 16) + (uint32_t)cb, v);
      NB_GCLK(9)
      if (a.bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += sbias[cb + i];
      }
      const int sw = row & 15;
      if (a.out_pre) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          nb_st4(stage + row * NB_H + ((((cb >> 2) + k) ^ sw) << 2), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int idx = tid + it * NB_THREADS, r = idx >> 4, ch = idx & 15;
          if (r < nv) nb_st4(a.out_pre + (int64_t)(r0 + r) * a.ldp + 4 * ch, nb_ld4(stage + r * NB_H + ((ch ^ (r & 15)) << 2)));
        }
        __syncthreads();
      }
      if (a.out) {
        if (live) {
          if (a.epi == NB_EPI_SILU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = nb_silu(v[i]);
          } else if (a.epi == NB_EPI_MUL_DSILU) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 u = NSRC == 1 ? eu[NSRC == 1 ? k : 0] : nb_ld4(a.U + gr * a.ldu + cb + 4 * k);
              v[4 * k + 0] *= nb_dsilu(u.x);
              v[4 * k + 1] *= nb_dsilu(u.y);
              v[4 * k + 2] *= nb_dsilu(u.z);
              v[4 * k + 3] *= nb_dsilu(u.w);
            }
          }
          if (a.R) {
            const float* rp = a.R + gr * a.ldr + cb;
            const bool pre = NSRC == 1 && a.epi != NB_EPI_MUL_DSILU;   // prefetched above unless U took the slot
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float4 t = pre ? eu[NSRC == 1 ? k : 0] : nb_ld4(rp + 4 * k);
              v[4 * k + 0] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
            }
          }
          if (a.accumulate) {
            const float* o = a.out + gr * a.ldo + cb;
            const bool pre = NSRC == 1 && a.epi != NB_EPI_MUL_DSILU && !a.R;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float4 t = pre ? eu[NSRC == 1 ? k : 0] : nb_ld4(o + 4 * k);
              v[4 * k + 0] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          nb_st4(stage + row * NB_H + ((((cb >> 2) + k) ^ sw) << 2), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int idx = tid + it * NB_THREADS, r = idx >> 4, ch = idx & 15;
          if (r < nv) nb_st4(a.out + (int64_t)(r0 + r) * a.ldo + 4 * ch, nb_ld4(stage + r * NB_H + ((ch ^ (r & 15)) << 2)));
        }
        __syncthreads();   // the next tile's split pieces overwrite the staging area
      }
    }
    nb_tc_fence_before();  // the next tile's MMA overwrites the accumulator this thread has just read
    NB_GCLK(10)
  }
  nb_tc_fence_before();
  __syncthreads();
  NB_GCLK(11)
  if (warp == 0) nb_tmem_dealloc(tm, 64);
  NB_GCLK(12)
}

__global__ void __launch_bounds__(NB_THREADS, 2) k_gemm64_tc(NbGemmBatch batch, int nsrc_max) {
  NB_PDL_ENTER();
  const NbGemmArgs& a = batch.job[blockIdx.y];
  const int ntiles = (a.rows + NB_TILE - 1) / NB_TILE;
  if ((int)blockIdx.x >= ntiles) return;
  extern __shared__ __align__(1024) unsigned char nb_smraw[];
  unsigned char* base = nb_smraw + ((1024u - (nb_smem_u32(nb_smraw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + NB_GT_A(nsrc_max));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  if (a.nsrc <= 1) nb_gemm64_tc_body<1>(a, base, bar, tmem_slot, nsrc_max);   // block-uniform: one job per blockIdx.y
  else nb_gemm64_tc_body<2>(a, base, bar, tmem_slot, nsrc_max);
}

// ============================================================================= weight images
// One CTA per 64 x 64 weight block W[r * ld + c]: split into bf16 pieces and written to global memory in the shared-memory
// tile layout (hi tile | lo tile), tile row = r, tile column = c.  K-major view: B[k = c][n = r] (y = x W^T);
// MN-major view: B[k = r][n = c] (g_in = g_out W).
#define NB_MAX_WIMG 64
struct NbWimgBatch {
  const float* W[NB_MAX_WIMG];
  int ld[NB_MAX_WIMG];
  unsigned char* out;  // image i at out + i * 2 * NB_TC_TILE_BYTES(64)
};
__global__ void __launch_bounds__(256) k_weight_images(NbWimgBatch b) {
  NB_PDL_ENTER();
  const float* W = b.W[blockIdx.x];
  const int ld = b.ld[blockIdx.x];
  unsigned char* hi = b.out + (size_t)blockIdx.x * 2 * NB_TC_TILE_BYTES(64);
  unsigned char* lo = hi + NB_TC_TILE_BYTES(64);
  for (int idx = threadIdx.x; idx < 64 * 8; idx += blockDim.x) {
    const int r = idx >> 3, j = idx & 7;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(W + (int64_t)r * ld + 8 * j + i);
    nb_tc_store8(hi, lo, r, j, v);
  }
}

// ============================================================================= wgrad64 on tensor cores
// partial[cta][o*64 + k] = sum over this CTA's rows of sum_p scale_p G_p[r][o] act_p(A_p[r][k]) ; [4096 + o] = colsum(G_0)
// D[64 x 64] = G^T A with both tiles MN-major (K = 128 rows), accumulated in TMEM over all tiles of the CTA.
#define NB_WT_SMEM (4 * NB_TC_TILE_BYTES(128) + NB_TILE * 16 + 64 + 1024)

__global__ void __launch_bounds__(NB_THREADS, 3) k_wgrad64_tc(NbWgradBatch batch) {
  NB_PDL_ENTER();
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
      // global loads first (registers), THEN the wait for the MMAs that still read the shared-memory tiles: the HBM
      // round trip of this pair hides under the previous pair's MMAs
      float4 g0[4], g1[4], a0[4], a1[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = tid + it * NB_THREADS;
        const int r = idx >> 3, j = idx & 7;
        if (r < nv) {
          const float* gp = nb_wg_row(pr.G, pr.ldg, pr.seg_rows, pr.seg_g, r0 + r) + 8 * j;
          const float* ap = nb_wg_row(pr.A, pr.lda, pr.seg_rows, pr.seg_a, r0 + r) + 8 * j;
          g0[it] = nb_ld4(gp); g1[it] = nb_ld4(gp + 4); a0[it] = nb_ld4(ap); a1[it] = nb_ld4(ap + 4);
        } else {
          g0[it] = g1[it] = a0[it] = a1[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (wacc) {  // the MMAs reading the tiles have completed
        nb_mbar_wait(bar, phase);
        phase ^= 1;
        nb_tc_fence_after();
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = tid + it * NB_THREADS;
        const int r = idx >> 3, j = idx & 7;
        float gv[8] = {g0[it].x, g0[it].y, g0[it].z, g0[it].w, g1[it].x, g1[it].y, g1[it].z, g1[it].w};
        float av[8] = {a0[it].x, a0[it].y, a0[it].z, a0[it].w, a1[it].x, a1[it].y, a1[it].z, a1[it].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) gv[i] *= pr.scale;
        if (pr.a_silu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) av[i] = nb_silu(av[i]);   // SiLU(0) = 0: padded rows stay zero
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
