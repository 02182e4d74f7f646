// nb_merge.cuh — multi-input SEGNO between its integration segments (SEGNO/models/model.py:65-90, 105-139).
//
// The reference embeds every observed frame (h = embedding(his), [BN, L, 64]), integrates from frame i to frame i + 1
// (forward_step) and merges the integrated state (x_i, h_i, v_i) with the observed frame i + 1 before the next segment:
//   'sum'  : h_ = h[:, i+1] + h_i ; x_ = x[:, i+1] + x_i ; v_ = v[:, i+1] + v_i                     (model.py:82-85)
//   'attn' : InvariantTemporalAttention over the pair (observed, integrated)                         (model.py:86-90, 105-139)
//            feats_c = [ |v_c| , h_c ]   a_c = tanh(W0 feats_c + b0)   l_c = w2 . a_c + b2   alpha = softmax_c(l)
//            (x_, v_, h_) = sum_c alpha_c (x_c, v_c, h_c)
// One warp owns a node: lane l owns hidden units l and l + 32; the 64 x 65 first-layer weights sit in shared memory with
// a row stride of 65 floats (row-wise reads by the lanes of the forward, column-wise reads of the backward's W0^T g are
// both conflict-free).  The backward recomputes a_c, keeps its 2 x 65 weight-gradient rows in registers over all the
// nodes of the warp, reduces the 8 warps of a CTA in a fixed order and leaves one partial slice per CTA for k_finalize:
// no atomics, bitwise deterministic.
#pragma once
#include "nb_common.cuh"

#define NB_AT_IN 65                                  // |v| + 64 hidden features
#define NB_AT_PLEN (NB_H * NB_AT_IN + 2 * NB_H + 1)  // W0[64][65] | b0[64] | w2[64] | b2  (named_parameters order)
#define NB_MERGE_COPY 0
#define NB_MERGE_SUM 1
#define NB_MERGE_ATTN 2

struct NbMergeArgs {
  int mode;
  int n, L, frame;                       // nodes; frames per node of the *_all tensors; the observed frame
  const float *h_all, *x_all, *v_all;    // [n][L][64], [n][L][3], [n][L][3]
  const float *h_int, *x_int, *v_int;    // integrated state [n][64], [n][3], [n][3]   (null for NB_MERGE_COPY)
  const float* ap;                       // attention parameters (NB_AT_PLEN floats)
  float *h_out, *x_out, *v_out;          // merged state [n][64], [n][3], [n][3]
  float* alpha;                          // [n][2]: written by the forward, read by the backward
  // backward
  const float *g_h, *g_x, *g_v;          // gradients of the merged state (null: zero)
  float *g_h_all, *g_x_all, *g_v_all;    // [n][L][.]: slice [:, frame, :] is written
  float *g_h_int, *g_x_int, *g_v_int;
  float* partial;                        // [grid][NB_AT_PLEN]
};

#define NB_MERGE_SMEM ((NB_H * NB_AT_IN + 2 * NB_H + 8 * 4 * NB_H) * sizeof(float))

// z_c[o] for this lane's two hidden units (o = lane, lane + 32) and both candidates; hb = [2][64] hidden rows of the warp
__device__ __forceinline__ void nb_attn_hidden(const float* W0s, const float* b0s, const float* hb, float s0, float s1,
                                               int lane, float (&a0)[2], float (&a1)[2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float* w = W0s + (lane + 32 * r) * NB_AT_IN;
    float z0 = fmaf(w[0], s0, b0s[lane + 32 * r]), z1 = fmaf(w[0], s1, b0s[lane + 32 * r]);
    for (int k = 0; k < NB_H; ++k) {
      z0 = fmaf(w[1 + k], hb[k], z0);
      z1 = fmaf(w[1 + k], hb[NB_H + k], z1);
    }
    a0[r] = tanhf(z0);
    a1[r] = tanhf(z1);
  }
}
__device__ __forceinline__ float nb_warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__global__ void __launch_bounds__(256) k_segno_merge_fwd(NbMergeArgs a) {
  NB_PDL_ENTER();
  NB_DYN_SMEM(sm);
  float* W0s = sm;                          // [64][65]
  float* b0s = W0s + NB_H * NB_AT_IN;       // [64]
  float* w2s = b0s + NB_H;                  // [64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* hb = w2s + NB_H + warp * 4 * NB_H;  // [2][64] of this warp (+ [2][64] used by the backward)
  float b2 = 0.f;
  if (a.mode == NB_MERGE_ATTN) {
    for (int i = threadIdx.x; i < NB_H * NB_AT_IN + 2 * NB_H; i += 256) sm[i] = __ldg(a.ap + i);
    b2 = __ldg(a.ap + NB_AT_PLEN - 1);
  }
  __syncthreads();
  for (int node = blockIdx.x * 8 + warp; node < a.n; node += gridDim.x * 8) {
    const int64_t ro = (int64_t)node * a.L + a.frame;
    const float2 ho = make_float2(a.h_all[ro * NB_H + lane], a.h_all[ro * NB_H + lane + 32]);
    const float xo = lane < 3 ? a.x_all[ro * 3 + lane] : 0.f, vo = lane < 3 ? a.v_all[ro * 3 + lane] : 0.f;
    float2 hi = make_float2(0.f, 0.f);
    float xi = 0.f, vi = 0.f;
    if (a.mode != NB_MERGE_COPY) {
      hi = make_float2(a.h_int[(int64_t)node * NB_H + lane], a.h_int[(int64_t)node * NB_H + lane + 32]);
      if (lane < 3) {
        xi = a.x_int[(int64_t)node * 3 + lane];
        vi = a.v_int[(int64_t)node * 3 + lane];
      }
    }
    float al0 = 1.f, al1 = a.mode == NB_MERGE_SUM ? 1.f : 0.f;
    if (a.mode == NB_MERGE_ATTN) {
      __syncwarp();
      hb[lane] = ho.x; hb[lane + 32] = ho.y; hb[NB_H + lane] = hi.x; hb[NB_H + lane + 32] = hi.y;
      __syncwarp();
      const float s0 = sqrtf(nb_warp_sum(vo * vo)), s1 = sqrtf(nb_warp_sum(vi * vi));
      float a0[2], a1[2];
      nb_attn_hidden(W0s, b0s, hb, s0, s1, lane, a0, a1);
      const float l0 = nb_warp_sum(w2s[lane] * a0[0] + w2s[lane + 32] * a0[1]) + b2;
      const float l1 = nb_warp_sum(w2s[lane] * a1[0] + w2s[lane + 32] * a1[1]) + b2;
      const float mx = fmaxf(l0, l1), e0 = expf(l0 - mx), e1 = expf(l1 - mx);
      al0 = e0 / (e0 + e1);
      al1 = e1 / (e0 + e1);
      if (lane == 0 && a.alpha) {
        a.alpha[2 * node] = al0;
        a.alpha[2 * node + 1] = al1;
      }
    }
    a.h_out[(int64_t)node * NB_H + lane] = al0 * ho.x + al1 * hi.x;
    a.h_out[(int64_t)node * NB_H + lane + 32] = al0 * ho.y + al1 * hi.y;
    if (lane < 3) {
      a.x_out[(int64_t)node * 3 + lane] = al0 * xo + al1 * xi;
      a.v_out[(int64_t)node * 3 + lane] = al0 * vo + al1 * vi;
    }
  }
}

__global__ void __launch_bounds__(256) k_segno_merge_bwd(NbMergeArgs a) {
  NB_PDL_ENTER();
  NB_DYN_SMEM(sm);
  float* W0s = sm;
  float* b0s = W0s + NB_H * NB_AT_IN;
  float* w2s = b0s + NB_H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* hb = w2s + NB_H + warp * 4 * NB_H;  // [2][64] hidden rows
  float* gzb = hb + 2 * NB_H;                // [2][64] dL/dz of both candidates
  const bool attn = a.mode == NB_MERGE_ATTN;
  float b2 = 0.f;
  if (attn) {
    for (int i = threadIdx.x; i < NB_H * NB_AT_IN + 2 * NB_H; i += 256) sm[i] = __ldg(a.ap + i);
    b2 = __ldg(a.ap + NB_AT_PLEN - 1);
  }
  (void)b2;
  __syncthreads();
  float gW[2][NB_AT_IN];   // rows lane, lane + 32 of dL/dW0
  float gb0[2] = {0.f, 0.f}, gw2[2] = {0.f, 0.f}, gb2 = 0.f;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int k = 0; k < NB_AT_IN; ++k) gW[r][k] = 0.f;
  }
  for (int node = blockIdx.x * 8 + warp; node < a.n; node += gridDim.x * 8) {
    const int64_t ro = (int64_t)node * a.L + a.frame;
    const float2 gh = a.g_h ? make_float2(a.g_h[(int64_t)node * NB_H + lane], a.g_h[(int64_t)node * NB_H + lane + 32]) : make_float2(0.f, 0.f);
    const float gx = (a.g_x && lane < 3) ? a.g_x[(int64_t)node * 3 + lane] : 0.f;
    const float gv = (a.g_v && lane < 3) ? a.g_v[(int64_t)node * 3 + lane] : 0.f;
    float al0 = 1.f, al1 = a.mode == NB_MERGE_SUM ? 1.f : 0.f;
    float2 eh0 = make_float2(0.f, 0.f), eh1 = eh0;   // attention path: extra dL/dh_c (through the features)
    float ev0 = 0.f, ev1 = 0.f;                      // extra dL/dv_c (through |v_c|)
    if (attn) {
      const float2 ho = make_float2(a.h_all[ro * NB_H + lane], a.h_all[ro * NB_H + lane + 32]);
      const float2 hi = make_float2(a.h_int[(int64_t)node * NB_H + lane], a.h_int[(int64_t)node * NB_H + lane + 32]);
      const float xo = lane < 3 ? a.x_all[ro * 3 + lane] : 0.f, vo = lane < 3 ? a.v_all[ro * 3 + lane] : 0.f;
      const float xi = lane < 3 ? a.x_int[(int64_t)node * 3 + lane] : 0.f, vi = lane < 3 ? a.v_int[(int64_t)node * 3 + lane] : 0.f;
      al0 = a.alpha[2 * node];
      al1 = a.alpha[2 * node + 1];
      __syncwarp();
      hb[lane] = ho.x; hb[lane + 32] = ho.y; hb[NB_H + lane] = hi.x; hb[NB_H + lane + 32] = hi.y;
      __syncwarp();
      const float s0 = sqrtf(nb_warp_sum(vo * vo)), s1 = sqrtf(nb_warp_sum(vi * vi));
      float a0[2], a1[2];
      nb_attn_hidden(W0s, b0s, hb, s0, s1, lane, a0, a1);
      // dL/dalpha_c = <g, candidate c>
      const float ga0 = nb_warp_sum(gh.x * ho.x + gh.y * ho.y + gx * xo + gv * vo);
      const float ga1 = nb_warp_sum(gh.x * hi.x + gh.y * hi.y + gx * xi + gv * vi);
      const float dot = al0 * ga0 + al1 * ga1;
      const float gl0 = al0 * (ga0 - dot), gl1 = al1 * (ga1 - dot);   // softmax backward
      gb2 += gl0 + gl1;
      float gz0[2], gz1[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float w2 = w2s[lane + 32 * r];
        gw2[r] += gl0 * a0[r] + gl1 * a1[r];
        gz0[r] = gl0 * w2 * (1.f - a0[r] * a0[r]);
        gz1[r] = gl1 * w2 * (1.f - a1[r] * a1[r]);
        gb0[r] += gz0[r] + gz1[r];
        gW[r][0] = fmaf(gz0[r], s0, fmaf(gz1[r], s1, gW[r][0]));
#pragma unroll
        for (int k = 0; k < NB_H; ++k) gW[r][1 + k] = fmaf(gz0[r], hb[k], fmaf(gz1[r], hb[NB_H + k], gW[r][1 + k]));
        gzb[lane + 32 * r] = gz0[r];
        gzb[NB_H + lane + 32 * r] = gz1[r];
      }
      __syncwarp();
      // dL/dfeats_c[j] = sum_o W0[o][j] gz_c[o]: this lane takes j = 1 + lane, 1 + lane + 32 (the hidden features) and,
      // redundantly on every lane, j = 0 (the speed)
      float f00 = 0.f, f01 = 0.f, f10 = 0.f, f11 = 0.f, fs0 = 0.f, fs1 = 0.f;
      for (int o = 0; o < NB_H; ++o) {
        const float* w = W0s + o * NB_AT_IN;
        const float z0 = gzb[o], z1 = gzb[NB_H + o];
        f00 = fmaf(w[1 + lane], z0, f00);
        f01 = fmaf(w[1 + lane + 32], z0, f01);
        f10 = fmaf(w[1 + lane], z1, f10);
        f11 = fmaf(w[1 + lane + 32], z1, f11);
        fs0 = fmaf(w[0], z0, fs0);
        fs1 = fmaf(w[0], z1, fs1);
      }
      eh0 = make_float2(f00, f01);
      eh1 = make_float2(f10, f11);
      ev0 = s0 > 0.f ? fs0 * vo / s0 : 0.f;   // d|v|/dv = v / |v| (0 at the origin, as torch.norm's backward)
      ev1 = s1 > 0.f ? fs1 * vi / s1 : 0.f;
    }
    a.g_h_all[ro * NB_H + lane] = al0 * gh.x + eh0.x;
    a.g_h_all[ro * NB_H + lane + 32] = al0 * gh.y + eh0.y;
    if (lane < 3) {
      a.g_x_all[ro * 3 + lane] = al0 * gx;
      a.g_v_all[ro * 3 + lane] = al0 * gv + ev0;
    }
    if (a.mode != NB_MERGE_COPY) {
      a.g_h_int[(int64_t)node * NB_H + lane] = al1 * gh.x + eh1.x;
      a.g_h_int[(int64_t)node * NB_H + lane + 32] = al1 * gh.y + eh1.y;
      if (lane < 3) {
        a.g_x_int[(int64_t)node * 3 + lane] = al1 * gx;
        a.g_v_int[(int64_t)node * 3 + lane] = al1 * gv + ev1;
      }
    }
  }
  if (!attn) return;
  // the 8 warps of the CTA add their rows into the (now free) weight copy in a fixed order
  __syncthreads();
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float* dst = W0s + (lane + 32 * r) * NB_AT_IN;
#pragma unroll
        for (int k = 0; k < NB_AT_IN; ++k) dst[k] = (w == 0 ? 0.f : dst[k]) + gW[r][k];
        b0s[lane + 32 * r] = (w == 0 ? 0.f : b0s[lane + 32 * r]) + gb0[r];
        w2s[lane + 32 * r] = (w == 0 ? 0.f : w2s[lane + 32 * r]) + gw2[r];
      }
      if (lane == 0) hb[0] = gb2;   // this warp's own slot: summed below
    }
    __syncthreads();
  }
  float* out = a.partial + (int64_t)blockIdx.x * NB_AT_PLEN;
  for (int i = threadIdx.x; i < NB_H * NB_AT_IN + 2 * NB_H; i += 256) out[i] = sm[i];
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += (w2s + NB_H + w * 4 * NB_H)[0];
    out[NB_AT_PLEN - 1] = s;
  }
}

// dst += src
__global__ void __launch_bounds__(256) k_accumulate(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  NB_PDL_ENTER();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
