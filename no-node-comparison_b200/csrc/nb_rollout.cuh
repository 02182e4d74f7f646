// nb_rollout.cuh — the callers' per-step host work moved onto the device:
//   * featurisation of a frame (prepare_inputs, EGNO/main_simulation_simple_no.py:326-338; SEGNO/train_nbody.py:
//     119-123, :228-233): per-node speed |v| (+ charge), per-graph mean position, per-edge (q_i q_j, |x_i - x_j|^2) in
//     the canonical edge order — no edge_index gather, one kernel;
//   * conserved energy of a frame (utils.py:126-144 charged, :175-195 gravity), one value per trajectory, so the
//     rollout loop never leaves the device (the reference synchronises and runs numpy once per emitted frame).
// One CTA per (frame, trajectory); all reductions are in a fixed order (bitwise deterministic).
#pragma once
#include "nb_common.cuh"

struct NbFeatArgs {
  int B, N, with_charge;       // with_charge: nodes = [|v|, q] (EGNO) else [|v|] (SEGNO his)
  const float* loc;            // [B*N][3]
  const float* vel;            // [B*N][3]
  const float* charges;        // [B*N]
  const float* edge_attr_o;    // [B*N*(N-1)] static edge attribute (q_i q_j of the dataset) or null: q_i q_j from charges
  float* nodes;                // [B*N][1 + with_charge]
  float* loc_mean;             // [B*N][3] or null
  float* edge_attr;            // [B*N*(N-1)][2]
};

__global__ void __launch_bounds__(128) k_nbody_features(NbFeatArgs a) {
  NB_PDL_ENTER();
  NB_DYN_SMEM(sm);             // [N][4]: x, y, z, q
  __shared__ float mean[3];
  const int N = a.N, tid = threadIdx.x;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const int64_t n0 = (int64_t)b * N;
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) {
      sm[i * 4 + 0] = a.loc[(n0 + i) * 3 + 0];
      sm[i * 4 + 1] = a.loc[(n0 + i) * 3 + 1];
      sm[i * 4 + 2] = a.loc[(n0 + i) * 3 + 2];
      sm[i * 4 + 3] = a.charges[n0 + i];
      const float vx = a.vel[(n0 + i) * 3 + 0], vy = a.vel[(n0 + i) * 3 + 1], vz = a.vel[(n0 + i) * 3 + 2];
      const float sp = sqrtf(vx * vx + vy * vy + vz * vz);
      if (a.with_charge) {
        a.nodes[(n0 + i) * 2 + 0] = sp;
        a.nodes[(n0 + i) * 2 + 1] = sm[i * 4 + 3];
      } else {
        a.nodes[n0 + i] = sp;
      }
    }
    __syncthreads();
    if (a.loc_mean) {
      if (tid < 3) {
        float s = 0.f;
        for (int i = 0; i < N; ++i) s += sm[i * 4 + tid];
        mean[tid] = s / (float)N;
      }
      __syncthreads();
      for (int idx = tid; idx < N * 3; idx += blockDim.x) a.loc_mean[n0 * 3 + idx] = mean[idx % 3];
    }
    const int EPG = N * (N - 1);
    const int64_t e0 = (int64_t)b * EPG;
    for (int e = tid; e < EPG; e += blockDim.x) {
      const int i = e / (N - 1), jj = e - i * (N - 1);
      const int j = jj + (jj >= i ? 1 : 0);
      const float dx = sm[i * 4 + 0] - sm[j * 4 + 0], dy = sm[i * 4 + 1] - sm[j * 4 + 1], dz = sm[i * 4 + 2] - sm[j * 4 + 2];
      const float qq = a.edge_attr_o ? a.edge_attr_o[e0 + e] : sm[i * 4 + 3] * sm[j * 4 + 3];
      a.edge_attr[(e0 + e) * 2 + 0] = qq;
      a.edge_attr[(e0 + e) * 2 + 1] = dx * dx + dy * dy + dz * dz;
    }
  }
}

struct NbEnergyArgs {
  int F, B, N, kind;           // kind 0: charged (K + sum_{i<j} q_i q_j / r), 1: gravity (sum m v^2 / 2 - G sum_{i<j} m_i m_j / r)
  float G;
  const float* loc;            // [F][B*N][3]
  const float* vel;            // [F][B*N][3]
  const float* charges;        // [B*N] charges or masses (shared by the frames)
  float* out;                  // [F][B]
};

__global__ void __launch_bounds__(128) k_nbody_energy(NbEnergyArgs a) {
  NB_PDL_ENTER();
  NB_DYN_SMEM(sm);             // [N][4] + [128] partials
  float* part = sm + a.N * 4;
  const int N = a.N, tid = threadIdx.x;
  const int64_t total = (int64_t)a.F * a.B;
  for (int64_t fb = blockIdx.x; fb < total; fb += gridDim.x) {
    const int b = (int)(fb % a.B);
    const int64_t n0 = fb * N;           // frame-major: node (f, b, i) at (f*B + b)*N + i
    __syncthreads();
    float acc = 0.f;
    for (int i = tid; i < N; i += blockDim.x) {
      sm[i * 4 + 0] = a.loc[(n0 + i) * 3 + 0];
      sm[i * 4 + 1] = a.loc[(n0 + i) * 3 + 1];
      sm[i * 4 + 2] = a.loc[(n0 + i) * 3 + 2];
      const float q = a.charges[(int64_t)b * N + i];
      sm[i * 4 + 3] = q;
      const float vx = a.vel[(n0 + i) * 3 + 0], vy = a.vel[(n0 + i) * 3 + 1], vz = a.vel[(n0 + i) * 3 + 2];
      const float v2 = vx * vx + vy * vy + vz * vz;
      acc += 0.5f * (a.kind == 1 ? q * v2 : v2);
    }
    __syncthreads();
    const int npair = N * (N - 1) / 2;
    for (int p = tid; p < npair; p += blockDim.x) {
      // pair index -> (i < j), row-major over the strict upper triangle
      int i = (int)((2.f * N - 1.f - sqrtf((2.f * N - 1.f) * (2.f * N - 1.f) - 8.f * (float)p)) * 0.5f);
      while (i * (2 * N - i - 1) / 2 > p) --i;
      while ((i + 1) * (2 * N - i - 2) / 2 <= p) ++i;
      const int j = p - i * (2 * N - i - 1) / 2 + i + 1;
      const float dx = sm[i * 4 + 0] - sm[j * 4 + 0], dy = sm[i * 4 + 1] - sm[j * 4 + 1], dz = sm[i * 4 + 2] - sm[j * 4 + 2];
      const float d = sqrtf(dx * dx + dy * dy + dz * dz);
      if (d > 0.f) {
        const float w = sm[i * 4 + 3] * sm[j * 4 + 3] / d;
        acc += a.kind == 1 ? -a.G * w : w;
      }
    }
    part[tid] = acc;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int k = 0; k < (int)blockDim.x; ++k) s += part[k];
      a.out[fb] = s;
    }
  }
}

// ----------------------------------------------------------------------------- trajectory MSE
// The callers' loss: MSELoss(reduction='none')(pred, target).mean((0, 1, 3)) -> losses[T], then losses.mean() or
// losses[0] (EGNO/main_simulation_simple_no.py:273-276; SEGNO/train_nbody.py:163-165 with T = 1), and its gradient
// with respect to the prediction, in two launches instead of ~10 element-wise / reduction kernels.  Deterministic:
// per-CTA partial sums in a fixed order, no atomics.
struct NbMseArgs {
  int T, layout, only_first, nchunk;
  int64_t R;                   // rows per frame (B*N)
  const float* pred;           // [T][R][3], frame-major (the models' output layout)
  const float* target;         // layout 0: [T][R][3] ; layout 1: [R][T][3] (the data loader's [B, N, T, 3])
  float* grad;                 // [T][R][3] or null: d loss / d pred
  float* partial;              // [T][nchunk]
  float* losses;               // [T]
  float* loss;                 // [1]
};

__global__ void __launch_bounds__(256) k_traj_mse(NbMseArgs a) {
  NB_PDL_ENTER();
  __shared__ float part[256];
  const int t = blockIdx.y, tid = threadIdx.x;
  const int64_t n = a.R * 3;   // elements of one frame
  const float gs = a.only_first ? (t == 0 ? 2.f / (float)n : 0.f) : 2.f / ((float)n * (float)a.T);
  float acc = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + tid; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / 3;
    const int k = (int)(e - 3 * r);
    const float p = a.pred[(int64_t)t * n + e];
    const float y = a.layout ? a.target[(r * a.T + t) * 3 + k] : a.target[(int64_t)t * n + e];
    const float d = p - y;
    acc = fmaf(d, d, acc);
    if (a.grad) a.grad[(int64_t)t * n + e] = gs * d;
  }
  part[tid] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {   // fixed-shape tree: bitwise reproducible
    if (tid < w) part[tid] += part[tid + w];
    __syncthreads();
  }
  if (tid == 0) a.partial[(int64_t)t * a.nchunk + blockIdx.x] = part[0];
}

__global__ void __launch_bounds__(32) k_traj_mse_fin(NbMseArgs a) {
  NB_PDL_ENTER();
  __shared__ float ls[NB_MAX_T];
  const int t = threadIdx.x;
  if (t < a.T) {
    float s = 0.f;
    for (int c = 0; c < a.nchunk; ++c) s += a.partial[(int64_t)t * a.nchunk + c];
    s /= (float)(a.R * 3);
    ls[t] = s;
    a.losses[t] = s;
  }
  __syncthreads();
  if (t == 0) {
    float s = 0.f;
    if (a.only_first) s = ls[0];
    else {
      for (int k = 0; k < a.T; ++k) s += ls[k];
      s /= (float)a.T;
    }
    a.loss[0] = s;
  }
}
