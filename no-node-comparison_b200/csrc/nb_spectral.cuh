// nb_spectral.cuh — EGNO's temporal spectral convolution (layer_no.py:80-178) without cuFFT.
//
// T <= 16, so rfft -> truncated complex mode mixing -> irfft collapses to a tiny real DFT held in
// registers.  For mode m with theta = 2 pi m / T and input x[s]:
//   C_m = sum_s x[s] cos(theta s),  S_m = sum_s x[s] sin(theta s)            (X_m = C_m - i S_m)
//   P_m = C_m Wr + S_m Wi,          Q_m = C_m Wi - S_m Wr                    (Y_m = P_m + i Q_m)
//   y[t] = P_0 / T + sum_{0<m<T/2} (2/T)(P_m cos(theta t) - Q_m sin(theta t)) + [m = T/2] P_m (-1)^t / T
// (the imaginary parts of the DC and Nyquist coefficients are discarded by irfft, so Im W_0 and
//  Im W_{T/2} never matter and receive zero gradient — SURVEY.md §3.2).
// The 64x64 channel mixes run through k_gemm64 on coefficient arrays; the kernels here do the
// per-trajectory DFT / inverse DFT, the LeakyReLU + residual, and the 2-channel (x - mean, v) variant.
#pragma once
#include "nb_common.cuh"

#define NB_MAX_MODES (NB_MAX_T / 2 + 1)
#define NB_MAX_COEF (2 * NB_MAX_MODES - 1)

struct NbTwiddle {
  int T, modes, ncoef;
  int nyq;  // index of the Nyquist mode (T even and modes > T/2), else -1
  float c[NB_MAX_MODES][NB_MAX_T];
  float s[NB_MAX_MODES][NB_MAX_T];
};

// coefficient slot of mode m: cosine/real part at ci(m), sine part at ci(m)+1 (absent for m = 0 and Nyquist)
__host__ __device__ __forceinline__ int nb_coef_index(const NbTwiddle& tw, int m) { return m == 0 ? 0 : 2 * m - 1; }

// ----------------------------------------------------------------------------- h path (64 channels)
struct NbDftArgs {
  NbTwiddle tw;
  int Nn0;            // node-trajectories
  const float* x;     // [T][Nn0][64]
  float* coef;        // [ncoef][Nn0][64]
  const float* ycoef; // [ncoef][Nn0][64]  (mixed coefficients P/Q)
  float* out;         // [T][Nn0][64]
  const float* gout;  // [T][Nn0][64]
  float* gycoef;      // [ncoef][Nn0][64]
  const float* gcoef; // [ncoef][Nn0][64]
  float* gx;          // [T][Nn0][64]
};

__device__ __forceinline__ float4 nb_f4_fma(float s, float4 a, float4 b) {
  return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w));
}

// coef = DFT(x)
__global__ void __launch_bounds__(256) k_dft_fwd(NbDftArgs a) {
  NB_PDL_ENTER();
  const int64_t total = (int64_t)a.Nn0 * 16;
  const int64_t plane = (int64_t)a.Nn0 * NB_H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = idx * 4;  // (k, c4) -> k*64 + c4*4
    float4 xs[NB_MAX_T];
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) xs[t] = nb_ld4(a.x + t * plane + off);
    for (int m = 0; m < a.tw.modes; ++m) {
      float4 C = make_float4(0.f, 0.f, 0.f, 0.f), S = C;
#pragma unroll
      for (int t = 0; t < NB_MAX_T; ++t)
        if (t < a.tw.T) {
          C = nb_f4_fma(a.tw.c[m][t], xs[t], C);
          S = nb_f4_fma(a.tw.s[m][t], xs[t], S);
        }
      int ci = nb_coef_index(a.tw, m);
      nb_st4(a.coef + ci * plane + off, C);
      if (m != 0 && m != a.tw.nyq) nb_st4(a.coef + (ci + 1) * plane + off, S);
    }
  }
}

// y[t] from mixed coefficients (helper shared by forward and backward)
__device__ __forceinline__ void nb_idft_eval(const NbTwiddle& tw, const float* ycoef, int64_t plane, int64_t off,
                                             float4 (&y)[NB_MAX_T]) {
  const float invT = 1.0f / (float)tw.T;
#pragma unroll
  for (int t = 0; t < NB_MAX_T; ++t) y[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int m = 0; m < tw.modes; ++m) {
    int ci = nb_coef_index(tw, m);
    float4 P = nb_ld4(ycoef + ci * plane + off);
    if (m == 0 || m == tw.nyq) {
#pragma unroll
      for (int t = 0; t < NB_MAX_T; ++t)
        if (t < tw.T) y[t] = nb_f4_fma(tw.c[m][t] * invT, P, y[t]);
    } else {
      float4 Q = nb_ld4(ycoef + (ci + 1) * plane + off);
#pragma unroll
      for (int t = 0; t < NB_MAX_T; ++t)
        if (t < tw.T) {
          y[t] = nb_f4_fma(2.f * invT * tw.c[m][t], P, y[t]);
          y[t] = nb_f4_fma(-2.f * invT * tw.s[m][t], Q, y[t]);
        }
    }
  }
}

__device__ __forceinline__ float nb_leaky(float v) { return v > 0.f ? v : 0.01f * v; }
__device__ __forceinline__ float nb_dleaky(float v) { return v > 0.f ? 1.f : 0.01f; }

// out = x + LeakyReLU(irfft(ycoef))        (TimeConv.forward, layer_no.py:121-126)
__global__ void __launch_bounds__(256) k_idft_fwd(NbDftArgs a) {
  NB_PDL_ENTER();
  const int64_t total = (int64_t)a.Nn0 * 16;
  const int64_t plane = (int64_t)a.Nn0 * NB_H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = idx * 4;
    float4 y[NB_MAX_T];
    nb_idft_eval(a.tw, a.ycoef, plane, off, y);
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) {
        float4 xv = nb_ld4(a.x + t * plane + off);
        nb_st4(a.out + t * plane + off, make_float4(xv.x + nb_leaky(y[t].x), xv.y + nb_leaky(y[t].y),
                                                    xv.z + nb_leaky(y[t].z), xv.w + nb_leaky(y[t].w)));
      }
  }
}

// gycoef = adjoint of the inverse DFT applied to gout * LeakyReLU'(y)
__global__ void __launch_bounds__(256) k_idft_bwd(NbDftArgs a) {
  NB_PDL_ENTER();
  const int64_t total = (int64_t)a.Nn0 * 16;
  const int64_t plane = (int64_t)a.Nn0 * NB_H;
  const float invT = 1.0f / (float)a.tw.T;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = idx * 4;
    float4 y[NB_MAX_T];
    nb_idft_eval(a.tw, a.ycoef, plane, off, y);
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) {
        float4 g = nb_ld4(a.gout + t * plane + off);
        y[t] = make_float4(g.x * nb_dleaky(y[t].x), g.y * nb_dleaky(y[t].y), g.z * nb_dleaky(y[t].z),
                           g.w * nb_dleaky(y[t].w));  // y now holds gy
      }
    for (int m = 0; m < a.tw.modes; ++m) {
      float4 gP = make_float4(0.f, 0.f, 0.f, 0.f), gQ = gP;
      const bool single = (m == 0 || m == a.tw.nyq);
      const float sc = single ? invT : 2.f * invT;
#pragma unroll
      for (int t = 0; t < NB_MAX_T; ++t)
        if (t < a.tw.T) {
          gP = nb_f4_fma(sc * a.tw.c[m][t], y[t], gP);
          gQ = nb_f4_fma(-sc * a.tw.s[m][t], y[t], gQ);
        }
      int ci = nb_coef_index(a.tw, m);
      nb_st4(a.gycoef + ci * plane + off, gP);
      if (!single) nb_st4(a.gycoef + (ci + 1) * plane + off, gQ);
    }
  }
}

// gx[s] = gout[s] + sum_m gC_m cos(theta s) + gS_m sin(theta s)
__global__ void __launch_bounds__(256) k_dft_bwd(NbDftArgs a) {
  NB_PDL_ENTER();
  const int64_t total = (int64_t)a.Nn0 * 16;
  const int64_t plane = (int64_t)a.Nn0 * NB_H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = idx * 4;
    float4 g[NB_MAX_T];
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) g[t] = nb_ld4(a.gout + t * plane + off);
    for (int m = 0; m < a.tw.modes; ++m) {
      int ci = nb_coef_index(a.tw, m);
      float4 gC = nb_ld4(a.gcoef + ci * plane + off);
      const bool single = (m == 0 || m == a.tw.nyq);
      float4 gS = single ? make_float4(0.f, 0.f, 0.f, 0.f) : nb_ld4(a.gcoef + (ci + 1) * plane + off);
#pragma unroll
      for (int t = 0; t < NB_MAX_T; ++t)
        if (t < a.tw.T) {
          g[t] = nb_f4_fma(a.tw.c[m][t], gC, g[t]);
          if (!single) g[t] = nb_f4_fma(a.tw.s[m][t], gS, g[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) nb_st4(a.gx + t * plane + off, g[t]);
  }
}

// ----------------------------------------------------------------------------- (x - mean, v) path, 2 channels
// X[t] = (x0[t] - mean, v0[t]);  Xo = X + conv(X);  x1 = Xo[0] + mean, v1 = Xo[1]   (egno.py:103-108)
// weights W[i][o][m][2] (i, o in {0,1}).  One thread per (node-trajectory, xyz component).
struct NbTcxArgs {
  NbTwiddle tw;
  int n3;              // Nn0 * 3
  const float* x0;     // [T][n3]
  const float* v0;     // [T][n3]
  const float* mean;   // [L][n3]: frame t uses mean[tmap[t]]  (single input: L = 1, tmap = 0)
  int tmap[NB_MAX_T];
  const float* W;      // [2][2][modes][2]
  float *x1, *v1;      // [T][n3]
  // backward
  const float *gx1, *gv1;
  float *gx0, *gv0;
  float* partial;      // [grid][2*2*modes*2]
};

__device__ __forceinline__ float nb_tcx_w(const NbTcxArgs& a, int i, int o, int m, int c) {
  return __ldg(a.W + ((i * 2 + o) * a.tw.modes + m) * 2 + c);
}

__global__ void __launch_bounds__(256) k_tcx_fwd(NbTcxArgs a) {
  NB_PDL_ENTER();
  const float invT = 1.0f / (float)a.tw.T;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < a.n3; idx += gridDim.x * blockDim.x) {
    float X[2][NB_MAX_T], Y[2][NB_MAX_T], mean[NB_MAX_T];
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) {
        mean[t] = a.mean[(int64_t)a.tmap[t] * a.n3 + idx];
        X[0][t] = a.x0[(int64_t)t * a.n3 + idx] - mean[t];
        X[1][t] = a.v0[(int64_t)t * a.n3 + idx];
        Y[0][t] = 0.f;
        Y[1][t] = 0.f;
      }
    for (int m = 0; m < a.tw.modes; ++m) {
      float C[2] = {0.f, 0.f}, S[2] = {0.f, 0.f};
#pragma unroll
      for (int t = 0; t < NB_MAX_T; ++t)
        if (t < a.tw.T) {
          C[0] = fmaf(a.tw.c[m][t], X[0][t], C[0]);
          C[1] = fmaf(a.tw.c[m][t], X[1][t], C[1]);
          S[0] = fmaf(a.tw.s[m][t], X[0][t], S[0]);
          S[1] = fmaf(a.tw.s[m][t], X[1][t], S[1]);
        }
      const bool single = (m == 0 || m == a.tw.nyq);
      const float sc = single ? invT : 2.f * invT;
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        float P = 0.f, Q = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float wr = nb_tcx_w(a, i, o, m, 0), wi = nb_tcx_w(a, i, o, m, 1);
          P += C[i] * wr + S[i] * wi;
          Q += C[i] * wi - S[i] * wr;
        }
#pragma unroll
        for (int t = 0; t < NB_MAX_T; ++t)
          if (t < a.tw.T) {
            Y[o][t] = fmaf(sc * a.tw.c[m][t], P, Y[o][t]);
            if (!single) Y[o][t] = fmaf(-sc * a.tw.s[m][t], Q, Y[o][t]);
          }
      }
    }
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) {
        a.x1[(int64_t)t * a.n3 + idx] = (X[0][t] + Y[0][t]) + mean[t];
        a.v1[(int64_t)t * a.n3 + idx] = X[1][t] + Y[1][t];
      }
  }
}

__global__ void __launch_bounds__(256) k_tcx_bwd(NbTcxArgs a) {
  NB_PDL_ENTER();
  __shared__ float red[8][2 * 2 * NB_MAX_MODES * 2];
  const float invT = 1.0f / (float)a.tw.T;
  const int nw = 2 * 2 * a.tw.modes * 2;
  float gw[2 * 2 * NB_MAX_MODES * 2];
#pragma unroll
  for (int i = 0; i < 2 * 2 * NB_MAX_MODES * 2; ++i) gw[i] = 0.f;
  for (int base = blockIdx.x * blockDim.x; base < a.n3; base += gridDim.x * blockDim.x) {
    const int idx = base + threadIdx.x;
    const bool act = idx < a.n3;
    float X[2][NB_MAX_T], G[2][NB_MAX_T], GX[2][NB_MAX_T];
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) {
        X[0][t] = act ? a.x0[(int64_t)t * a.n3 + idx] - a.mean[(int64_t)a.tmap[t] * a.n3 + idx] : 0.f;
        X[1][t] = act ? a.v0[(int64_t)t * a.n3 + idx] : 0.f;
        G[0][t] = act ? a.gx1[(int64_t)t * a.n3 + idx] : 0.f;
        G[1][t] = act ? a.gv1[(int64_t)t * a.n3 + idx] : 0.f;
        GX[0][t] = G[0][t];  // residual path
        GX[1][t] = G[1][t];
      }
    for (int m = 0; m < a.tw.modes; ++m) {
      float C[2] = {0.f, 0.f}, S[2] = {0.f, 0.f}, gP[2] = {0.f, 0.f}, gQ[2] = {0.f, 0.f};
      const bool single = (m == 0 || m == a.tw.nyq);
      const float sc = single ? invT : 2.f * invT;
#pragma unroll
      for (int t = 0; t < NB_MAX_T; ++t)
        if (t < a.tw.T) {
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            C[ch] = fmaf(a.tw.c[m][t], X[ch][t], C[ch]);
            S[ch] = fmaf(a.tw.s[m][t], X[ch][t], S[ch]);
            gP[ch] = fmaf(sc * a.tw.c[m][t], G[ch][t], gP[ch]);
            gQ[ch] = fmaf(-sc * a.tw.s[m][t], G[ch][t], gQ[ch]);
          }
        }
      if (single) gQ[0] = gQ[1] = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float gC = 0.f, gS = 0.f;
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          float wr = nb_tcx_w(a, i, o, m, 0), wi = nb_tcx_w(a, i, o, m, 1);
          gC += gP[o] * wr + gQ[o] * wi;
          gS += gP[o] * wi - gQ[o] * wr;
          // dL/dWr[i][o] = C_i gP_o - S_i gQ_o ;  dL/dWi[i][o] = S_i gP_o + C_i gQ_o
          gw[((i * 2 + o) * NB_MAX_MODES + m) * 2 + 0] += C[i] * gP[o] - S[i] * gQ[o];
          if (!single) gw[((i * 2 + o) * NB_MAX_MODES + m) * 2 + 1] += S[i] * gP[o] + C[i] * gQ[o];
        }
#pragma unroll
        for (int t = 0; t < NB_MAX_T; ++t)
          if (t < a.tw.T) {
            GX[i][t] = fmaf(a.tw.c[m][t], gC, GX[i][t]);
            if (!single) GX[i][t] = fmaf(a.tw.s[m][t], gS, GX[i][t]);
          }
      }
    }
    if (act) {
#pragma unroll
      for (int t = 0; t < NB_MAX_T; ++t)
        if (t < a.tw.T) {
          a.gx0[(int64_t)t * a.n3 + idx] = GX[0][t];
          a.gv0[(int64_t)t * a.n3 + idx] = GX[1][t];
        }
    }
  }
  // block reduction of the 8*modes weight gradients: warp shuffle, then across the 8 warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 2 * 2 * NB_MAX_MODES * 2; ++i) {
    int m = (i >> 1) % NB_MAX_MODES;
    if (m >= a.tw.modes) continue;
    float v = gw[i];
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < nw) {
    // compact index (i*2+o, m, c) with the real mode count
    int c = threadIdx.x & 1, m = (threadIdx.x >> 1) % a.tw.modes, io = (threadIdx.x >> 1) / a.tw.modes;
    int src = (io * NB_MAX_MODES + m) * 2 + c;
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][src];
    a.partial[(int64_t)blockIdx.x * nw + threadIdx.x] = s;
  }
}

// ----------------------------------------------------------------------------- fused temporal convolution (modes <= 2)
// One kernel for  out = x + LeakyReLU(irfft(mix(rfft(x))))  and one for its backward, for the configured case
// num_modes <= 2 (model_confs.yaml:1-13): coefficient planes C0 | C1 | S1.  A CTA processes groups of 16
// node-trajectories; thread (r = tid / 16, c4 = tid % 16) owns row r and channels 4 c4 .. 4 c4 + 3 in BOTH the DFT and
// the 64x64 mode-mixing GEMM (fp32 FFMA: the result feeds LeakyReLU, see launch_gemm_batch), so the mixed
// coefficients come out in the registers of the thread that holds x[t] — x and out cross HBM exactly once.
//   forward : 2 x T x 256 B per node-trajectory;  backward: reads x, gout, writes gx (+ coef, gycoef planes for the
//   weight-gradient reduction).
#define NB_TCV_ROWS 16
#define NB_TCV_LDA 68
struct NbTconvArgs {
  NbTwiddle tw;
  int Nn0;
  const float* x;      // [T][Nn0][64]
  const float* W;      // [64][64][modes][2]  (i, o, m, re/im)
  float* out;          // [T][Nn0][64]                 forward
  const float* gout;   // [T][Nn0][64]                 backward
  float* gx;           // [T][Nn0][64]
  float* coef;         // [ncoef][Nn0][64]  C0 | C1 | S1   (backward: written for the weight-gradient jobs)
  float* gycoef;       // [ncoef][Nn0][64]  gP0 | gP1 | gQ1
  // [Nn0][16][2] words: bit 4 t + c of word pair (row, c4) = [y_t[4 c4 + c] > 0], the LeakyReLU mask of the forward.
  // Training forward writes it, the backward reads it: no recompute of the mode mixing (and none of its rounding) there.
  uint32_t* mask;
};
#define NB_TCONV_FWD_SMEM ((3 * NB_H * NB_H + 3 * NB_TCV_ROWS * NB_TCV_LDA) * sizeof(float))
#define NB_TCONV_BWD_SMEM ((3 * NB_H * 68 + 3 * NB_TCV_ROWS * NB_TCV_LDA) * sizeof(float))

// De-interleave W[i][o][m][re/im] (modes <= 2) into the k-major planes B0 = Re W_0, B1 = Re W_1, B2 = Im W_1
// (Bq[i][o], row stride 64) and, when Tq != nullptr, their transposes Tq[o][i] (row stride NB_TCV_LDT).
// One coalesced pass: 8-byte loads (the parameter offset is only guaranteed to be 8-byte aligned).
#define NB_TCV_LDT 68
__device__ __forceinline__ void nb_tconv_stage_w(float* B0, float* B1, float* B2, float* T0, float* T1, float* T2,
                                                 const float* __restrict__ W, int modes, int tid) {
  // two batches of 8 iterations: 16 independent 8-byte loads in flight per thread before the first store
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float2 w0[8], w1[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = tid + (half * 8 + it) * 256;
      w0[it] = __ldg(reinterpret_cast<const float2*>(W + (int64_t)idx * modes * 2));
      w1[it] = modes > 1 ? __ldg(reinterpret_cast<const float2*>(W + (int64_t)idx * modes * 2 + 2)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = tid + (half * 8 + it) * 256;
      const int i = idx >> 6, o = idx & 63;
      if (B0) {
        B0[idx] = w0[it].x;
        B1[idx] = w1[it].x;
        B2[idx] = w1[it].y;
      }
      if (T0) {
        T0[o * NB_TCV_LDT + i] = w0[it].x;
        T1[o * NB_TCV_LDT + i] = w1[it].x;
        T2[o * NB_TCV_LDT + i] = w1[it].y;
      }
    }
  }
}

// acc_p (+)= sum_k A_p[row][k] * B_q[k][4 tx ..]  for the three coefficient planes:
//   out0 = A0 B0 ; out1 = A1 B1 + A2 B2 ; out2 = A1 B2 - A2 B1
__device__ __forceinline__ void nb_tconv_mix(const float* As, const float* B0, const float* B1, const float* B2, int ldb,
                                             int r, int c4, bool has1, bool pair1, float4& o0, float4& o1, float4& o2) {
  o0 = o1 = o2 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* a0 = As + r * NB_TCV_LDA;
  const float* a1 = a0 + NB_TCV_ROWS * NB_TCV_LDA;
  const float* a2 = a1 + NB_TCV_ROWS * NB_TCV_LDA;
#pragma unroll 2
  for (int k0 = 0; k0 < NB_H; k0 += 4) {
    const float4 v0 = nb_ld4(a0 + k0);
    const float4 v1 = has1 ? nb_ld4(a1 + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 v2 = pair1 ? nb_ld4(a2 + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float x0[4] = {v0.x, v0.y, v0.z, v0.w}, x1[4] = {v1.x, v1.y, v1.z, v1.w}, x2[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = k0 + kk;
      o0 = nb_f4_fma(x0[kk], nb_ld4(B0 + k * ldb + c4 * 4), o0);
      if (has1) {
        const float4 b1 = nb_ld4(B1 + k * ldb + c4 * 4);
        o1 = nb_f4_fma(x1[kk], b1, o1);
        if (pair1) {
          const float4 b2 = nb_ld4(B2 + k * ldb + c4 * 4);
          o1 = nb_f4_fma(x2[kk], b2, o1);
          o2 = nb_f4_fma(x1[kk], b2, o2);
          o2 = nb_f4_fma(-x2[kk], b1, o2);
        }
      }
    }
  }
}

// MAXT: compile-time bound of the frame loops (T <= MAXT): the per-frame register arrays are sized by it, so the
// configured T = 10 does not pay for 16 frames
template <int MAXT>
__global__ void __launch_bounds__(256, 3) k_tconv_fwd(NbTconvArgs a) {
  NB_PDL_ENTER();
  NB_DYN_SMEM(sm);
  float* B0 = sm;                    // W[.][.][0][re]
  float* B1 = B0 + NB_H * NB_H;      // W[.][.][1][re]
  float* B2 = B1 + NB_H * NB_H;      // W[.][.][1][im]
  float* As = B2 + NB_H * NB_H;      // [3][16][68]
  const int tid = threadIdx.x, r = tid >> 4, c4 = tid & 15;
  const NbTwiddle& tw = a.tw;
  const bool has1 = tw.modes > 1, pair1 = has1 && tw.nyq != 1;
  const int64_t plane = (int64_t)a.Nn0 * NB_H;
  const float invT = 1.0f / (float)tw.T;
  const int ngroups = (a.Nn0 + NB_TCV_ROWS - 1) / NB_TCV_ROWS;
  // the first group's rows are requested BEFORE the weight staging, so the two memory round trips overlap
  float4 xs[MAXT];
  {
    const int row = blockIdx.x * NB_TCV_ROWS + r;
    const bool act = (int)blockIdx.x < ngroups && row < a.Nn0;
    const int64_t off = (int64_t)row * NB_H + c4 * 4;
#pragma unroll
    for (int t = 0; t < MAXT; ++t)
      if (t < tw.T) xs[t] = act ? nb_ld4(a.x + t * plane + off) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  nb_tconv_stage_w(B0, B1, B2, nullptr, nullptr, nullptr, a.W, tw.modes, tid);
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int row = grp * NB_TCV_ROWS + r;
    const bool act = row < a.Nn0;
    const int64_t off = (int64_t)row * NB_H + c4 * 4;
    if (grp != (int)blockIdx.x) {
#pragma unroll
      for (int t = 0; t < MAXT; ++t)
        if (t < tw.T) xs[t] = act ? nb_ld4(a.x + t * plane + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 C0 = make_float4(0.f, 0.f, 0.f, 0.f), C1 = C0, S1 = C0;
#pragma unroll
    for (int t = 0; t < MAXT; ++t)
      if (t < tw.T) {
        C0 = nb_f4_fma(tw.c[0][t], xs[t], C0);
        if (has1) C1 = nb_f4_fma(tw.c[1][t], xs[t], C1);
        if (pair1) S1 = nb_f4_fma(tw.s[1][t], xs[t], S1);
      }
    __syncthreads();  // previous group's GEMM has finished reading As (also orders the weight staging)
    nb_st4(As + r * NB_TCV_LDA + c4 * 4, C0);
    nb_st4(As + (NB_TCV_ROWS + r) * NB_TCV_LDA + c4 * 4, C1);
    nb_st4(As + (2 * NB_TCV_ROWS + r) * NB_TCV_LDA + c4 * 4, S1);
    __syncthreads();
    float4 P0, P1, Q1;
    nb_tconv_mix(As, B0, B1, B2, NB_H, r, c4, has1, pair1, P0, P1, Q1);
    if (act) {
      const float s1 = pair1 ? 2.f * invT : invT;
      uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
      for (int t = 0; t < MAXT; ++t)
        if (t < tw.T) {
          float4 y = make_float4(P0.x * invT, P0.y * invT, P0.z * invT, P0.w * invT);
          if (has1) y = nb_f4_fma(s1 * tw.c[1][t], P1, y);
          if (pair1) y = nb_f4_fma(-s1 * tw.s[1][t], Q1, y);
          nb_st4(a.out + t * plane + off, make_float4(xs[t].x + nb_leaky(y.x), xs[t].y + nb_leaky(y.y),
                                                      xs[t].z + nb_leaky(y.z), xs[t].w + nb_leaky(y.w)));
          const uint32_t bits = (y.x > 0.f ? 1u : 0u) | (y.y > 0.f ? 2u : 0u) | (y.z > 0.f ? 4u : 0u) | (y.w > 0.f ? 8u : 0u);
          if (t < 8) m0 |= bits << (4 * t);
          else m1 |= bits << (4 * (t - 8));
        }
      if (a.mask) {
        a.mask[((int64_t)row * 16 + c4) * 2] = m0;
        a.mask[((int64_t)row * 16 + c4) * 2 + 1] = m1;
      }
    }
  }
}

template <int MAXT>
__global__ void __launch_bounds__(256, 2) k_tconv_bwd(NbTconvArgs a) {
  NB_PDL_ENTER();
  NB_DYN_SMEM(sm);
  float* T0 = sm;                      // transposed weight planes for the data gradient, row stride NB_TCV_LDT
  float* T1 = T0 + NB_H * NB_TCV_LDT;
  float* T2 = T1 + NB_H * NB_TCV_LDT;
  float* As = T2 + NB_H * NB_TCV_LDT;  // [3][16][68]
  const int tid = threadIdx.x, r = tid >> 4, c4 = tid & 15;
  const NbTwiddle& tw = a.tw;
  const bool has1 = tw.modes > 1, pair1 = has1 && tw.nyq != 1;
  const int64_t plane = (int64_t)a.Nn0 * NB_H;
  const float invT = 1.0f / (float)tw.T;
  const float s1 = pair1 ? 2.f * invT : invT;
  const int ngroups = (a.Nn0 + NB_TCV_ROWS - 1) / NB_TCV_ROWS;
  // The LeakyReLU mask comes from the forward (a.mask), so nothing of the forward mixing is recomputed here: one pass over
  // the frames gives the coefficients of x (for the weight gradients) and the adjoint of the inverse DFT applied to
  // gout * LeakyReLU'(y); ONE mixing GEMM (transposed weights) turns that into the coefficient gradients.  The first
  // group's gout rows are requested before the weight staging and stay in registers for the residual path.
  float4 gq[MAXT];
  {
    const int row = blockIdx.x * NB_TCV_ROWS + r;
    const bool act = (int)blockIdx.x < ngroups && row < a.Nn0;
    const int64_t off = (int64_t)row * NB_H + c4 * 4;
#pragma unroll
    for (int t = 0; t < MAXT; ++t)
      if (t < tw.T) gq[t] = act ? nb_ld4(a.gout + t * plane + off) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  nb_tconv_stage_w(nullptr, nullptr, nullptr, T0, T1, T2, a.W, tw.modes, tid);
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int row = grp * NB_TCV_ROWS + r;
    const bool act = row < a.Nn0;
    const int64_t off = (int64_t)row * NB_H + c4 * 4;
    if (grp != (int)blockIdx.x) {
#pragma unroll
      for (int t = 0; t < MAXT; ++t)
        if (t < tw.T) gq[t] = act ? nb_ld4(a.gout + t * plane + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint32_t mk0 = act ? a.mask[((int64_t)row * 16 + c4) * 2] : 0u, mk1 = act ? a.mask[((int64_t)row * 16 + c4) * 2 + 1] : 0u;
    float4 C0 = make_float4(0.f, 0.f, 0.f, 0.f), C1 = C0, S1 = C0, gP0 = C0, gP1 = C0, gQ1 = C0;
#pragma unroll
    for (int t = 0; t < MAXT; ++t)
      if (t < tw.T) {
        const float4 xv = act ? nb_ld4(a.x + t * plane + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        C0 = nb_f4_fma(tw.c[0][t], xv, C0);
        if (has1) C1 = nb_f4_fma(tw.c[1][t], xv, C1);
        if (pair1) S1 = nb_f4_fma(tw.s[1][t], xv, S1);
        const uint32_t bits = (t < 8 ? mk0 >> (4 * t) : mk1 >> (4 * (t - 8))) & 15u;
        const float4 g = gq[t];
        const float4 gy = make_float4(g.x * ((bits & 1u) ? 1.f : 0.01f), g.y * ((bits & 2u) ? 1.f : 0.01f),
                                      g.z * ((bits & 4u) ? 1.f : 0.01f), g.w * ((bits & 8u) ? 1.f : 0.01f));
        gP0 = nb_f4_fma(invT * tw.c[0][t], gy, gP0);
        if (has1) gP1 = nb_f4_fma(s1 * tw.c[1][t], gy, gP1);
        if (pair1) gQ1 = nb_f4_fma(-s1 * tw.s[1][t], gy, gQ1);
      }
    if (act) {
      nb_st4(a.coef + off, C0);
      if (has1) nb_st4(a.coef + plane + off, C1);
      if (pair1) nb_st4(a.coef + 2 * plane + off, S1);
      nb_st4(a.gycoef + off, gP0);
      if (has1) nb_st4(a.gycoef + plane + off, gP1);
      if (pair1) nb_st4(a.gycoef + 2 * plane + off, gQ1);
    }
    __syncthreads();  // the previous group's mix has finished reading As (also orders the weight staging)
    nb_st4(As + r * NB_TCV_LDA + c4 * 4, gP0);
    nb_st4(As + (NB_TCV_ROWS + r) * NB_TCV_LDA + c4 * 4, gP1);
    nb_st4(As + (2 * NB_TCV_ROWS + r) * NB_TCV_LDA + c4 * 4, gQ1);
    __syncthreads();
    // gC0 = gP0 Wr0^T ; gC1 = gP1 Wr1^T + gQ1 Wi1^T ; gS1 = gP1 Wi1^T - gQ1 Wr1^T   (same pattern as the forward mix)
    float4 gC0, gC1, gS1;
    nb_tconv_mix(As, T0, T1, T2, NB_TCV_LDT, r, c4, has1, pair1, gC0, gC1, gS1);
    if (act) {
#pragma unroll
      for (int t = 0; t < MAXT; ++t)
        if (t < tw.T) {
          float4 g = gq[t];  // residual path
          g = nb_f4_fma(tw.c[0][t], gC0, g);
          if (has1) g = nb_f4_fma(tw.c[1][t], gC1, g);
          if (pair1) g = nb_f4_fma(tw.s[1][t], gS1, g);
          nb_st4(a.gx + t * plane + off, g);
        }
    }
  }
}
