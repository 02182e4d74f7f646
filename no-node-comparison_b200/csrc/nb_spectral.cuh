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
  const float* mean;   // [n3]
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
  const float invT = 1.0f / (float)a.tw.T;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < a.n3; idx += gridDim.x * blockDim.x) {
    const float mean = a.mean[idx];
    float X[2][NB_MAX_T], Y[2][NB_MAX_T];
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) {
        X[0][t] = a.x0[(int64_t)t * a.n3 + idx] - mean;
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
        a.x1[(int64_t)t * a.n3 + idx] = (X[0][t] + Y[0][t]) + mean;
        a.v1[(int64_t)t * a.n3 + idx] = X[1][t] + Y[1][t];
      }
  }
}

__global__ void __launch_bounds__(256) k_tcx_bwd(NbTcxArgs a) {
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
    const float mean = act ? a.mean[idx] : 0.f;
#pragma unroll
    for (int t = 0; t < NB_MAX_T; ++t)
      if (t < a.tw.T) {
        X[0][t] = act ? a.x0[(int64_t)t * a.n3 + idx] - mean : 0.f;
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
