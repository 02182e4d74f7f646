// nb_sim.cuh — the reference's trajectory simulators on the device (last item): float64 leapfrog
// integrators of the charged and the gravitational N-body systems, one CTA per trajectory, thread i owns particle i.
//   ChargedParticlesSim.sample_trajectory  synthetic_sim.py:220-296 (+ _clamp :192-218, _l2 :165-177)
//   GravitySim.sample_trajectory           synthetic_sim.py:360-405 (+ compute_acceleration :311-333)
// Random draws (initial conditions, observation noise) stay on the host, in the reference's order; the kernels
// integrate.  The Python loops cost ~1 s per 100-body trajectory; here all trajectories of a data set advance together.
#pragma once
#include "nb_common.cuh"

#define NB_SIM_MAX_N 128

struct NbSimChargedArgs {
  int B, N, T, sample_freq;
  double dt, strength, max_f, box;
  const double* loc0;     // [B][3][N]
  const double* vel0;     // [B][3][N]
  const double* charges;  // [B][N]
  double* loc;            // [B][T / sample_freq - 1][3][N]
  double* vel;            // same (leapfrog half-step velocities, like the reference stores them)
};

// F_i = sum_{j != i} strength q_i q_j (x_i - x_j) / |x_i - x_j|^3, |.|^2 expanded as |a|^2 + |b|^2 - 2 a.b (the
// reference's _l2), components clipped to +-max_f
__device__ __forceinline__ void nb_sim_charged_force(const NbSimChargedArgs& a, double (*sx)[NB_SIM_MAX_N], double* sn,
                                                     const double* sq, int i, bool act, const double (&x)[3], double qi,
                                                     double (&F)[3]) {
  __syncthreads();  // the previous evaluation's readers are done
  if (act) {
    sx[0][i] = x[0];
    sx[1][i] = x[1];
    sx[2][i] = x[2];
    sn[i] = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  }
  __syncthreads();
  F[0] = F[1] = F[2] = 0.0;
  if (act) {
    for (int j = 0; j < a.N; ++j) {
      if (j == i) continue;
      const double dot = x[0] * sx[0][j] + x[1] * sx[1][j] + x[2] * sx[2][j];
      const double d2 = (sn[i] + sn[j]) - 2.0 * dot;
      const double w = a.strength * (qi * sq[j]) / pow(d2, 1.5);
      F[0] += w * (x[0] - sx[0][j]);
      F[1] += w * (x[1] - sx[1][j]);
      F[2] += w * (x[2] - sx[2][j]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) F[d] = fmin(fmax(F[d], -a.max_f), a.max_f);
  }
}

__global__ void __launch_bounds__(NB_SIM_MAX_N) k_sim_charged(NbSimChargedArgs a) {
  NB_PDL_ENTER();
  __shared__ double sx[3][NB_SIM_MAX_N];
  __shared__ double sn[NB_SIM_MAX_N];
  __shared__ double sq[NB_SIM_MAX_N];
  const int i = threadIdx.x, N = a.N;
  const bool act = i < N;
  const int nsave = a.T / a.sample_freq - 1;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    double x[3] = {0.0, 0.0, 0.0}, v[3] = {0.0, 0.0, 0.0}, F[3], qi = 0.0;
    __syncthreads();
    if (act) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        x[d] = a.loc0[((int64_t)b * 3 + d) * N + i];
        v[d] = a.vel0[((int64_t)b * 3 + d) * N + i];
        // _clamp: elastic reflection at the box walls, applied to the state the integration starts from
        if (x[d] > a.box) {
          x[d] = 2.0 * a.box - x[d];
          v[d] = -fabs(v[d]);
        }
        if (x[d] < -a.box) {
          x[d] = -2.0 * a.box - x[d];
          v[d] = fabs(v[d]);
        }
      }
      qi = a.charges[(int64_t)b * N + i];
      sq[i] = qi;
    }
    nb_sim_charged_force(a, sx, sn, sq, i, act, x, qi, F);
#pragma unroll
    for (int d = 0; d < 3; ++d) v[d] += a.dt * F[d];  // half step
    int k = 0;
    for (int it = 1; it < a.T; ++it) {
#pragma unroll
      for (int d = 0; d < 3; ++d) x[d] += a.dt * v[d];
      if (it % a.sample_freq == 0) {
        if (act && k < nsave) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            a.loc[(((int64_t)b * nsave + k) * 3 + d) * N + i] = x[d];
            a.vel[(((int64_t)b * nsave + k) * 3 + d) * N + i] = v[d];
          }
        }
        ++k;
      }
      nb_sim_charged_force(a, sx, sn, sq, i, act, x, qi, F);
#pragma unroll
      for (int d = 0; d < 3; ++d) v[d] += a.dt * F[d];
    }
  }
}

struct NbSimGravityArgs {
  int B, N, T, sample_freq;
  double dt, G, soft2;
  const double* pos0;  // [B][N][3]
  const double* vel0;  // [B][N][3]
  const double* mass;  // [B][N]
  double* pos;         // [B][T / sample_freq][N][3]
  double* vel;
  double* force;       // acc * mass
};

// a_i = G sum_j m_j (x_j - x_i) (|x_j - x_i|^2 + softening^2)^(-3/2)   (the j = i term is zero)
__device__ __forceinline__ void nb_sim_gravity_acc(const NbSimGravityArgs& a, double (*sx)[NB_SIM_MAX_N], const double* sm, int i,
                                                   bool act, const double (&x)[3], double (&acc)[3]) {
  __syncthreads();
  if (act) {
    sx[0][i] = x[0];
    sx[1][i] = x[1];
    sx[2][i] = x[2];
  }
  __syncthreads();
  acc[0] = acc[1] = acc[2] = 0.0;
  if (act) {
    for (int j = 0; j < a.N; ++j) {
      const double dx = sx[0][j] - x[0], dy = sx[1][j] - x[1], dz = sx[2][j] - x[2];
      const double w = pow(dx * dx + dy * dy + dz * dz + a.soft2, -1.5) * sm[j];
      acc[0] += dx * w;
      acc[1] += dy * w;
      acc[2] += dz * w;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) acc[d] *= a.G;
  }
}

__global__ void __launch_bounds__(NB_SIM_MAX_N) k_sim_gravity(NbSimGravityArgs a) {
  NB_PDL_ENTER();
  __shared__ double sx[3][NB_SIM_MAX_N];
  __shared__ double sm[NB_SIM_MAX_N];
  const int i = threadIdx.x, N = a.N;
  const bool act = i < N;
  const int nsave = a.T / a.sample_freq;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    double x[3] = {0.0, 0.0, 0.0}, v[3] = {0.0, 0.0, 0.0}, acc[3], mi = 0.0;
    __syncthreads();
    if (act) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        x[d] = a.pos0[((int64_t)b * N + i) * 3 + d];
        v[d] = a.vel0[((int64_t)b * N + i) * 3 + d];
      }
      mi = a.mass[(int64_t)b * N + i];
      sm[i] = mi;
    }
    nb_sim_gravity_acc(a, sx, sm, i, act, x, acc);
    for (int it = 0; it < a.T; ++it) {
      if (it % a.sample_freq == 0 && act) {
        const int64_t o = (((int64_t)b * nsave + it / a.sample_freq) * N + i) * 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          a.pos[o + d] = x[d];
          a.vel[o + d] = v[d];
          a.force[o + d] = acc[d] * mi;
        }
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        v[d] += acc[d] * a.dt / 2.0;  // kick
        x[d] += v[d] * a.dt;          // drift
      }
      nb_sim_gravity_acc(a, sx, sm, i, act, x, acc);
#pragma unroll
      for (int d = 0; d < 3; ++d) v[d] += acc[d] * a.dt / 2.0;  // kick
    }
  }
}
