"""CPU restatement (numpy, float64) of the reference's trajectory simulators — TEST INFRASTRUCTURE (the checker of the
CUDA simulator kernels, never the product path).

    ChargedParticlesSim.sample_trajectory   /root/reference/synthetic_sim.py:220-296  (+ _clamp :192-218, _l2 :165-177)
    GravitySim.sample_trajectory            /root/reference/synthetic_sim.py:360-405  (+ compute_acceleration :311-333)

Parity pin: tests/golden/sim_*.npz hold trajectories produced by the reference itself (tests/golden/make_sim_golden.py,
np.random.seed(43)) together with the initial conditions `charged_initial_conditions` / `gravity_initial_conditions`
draw from the same random stream; tests/test_sim.py holds this restatement to those vectors.
"""
from __future__ import annotations

import numpy as np

BOX = 5.0          # ChargedParticlesSim.box_size default (:150)
DT_CHARGED = 1e-3  # _delta_T (:162)


def charged_initial_conditions(n_balls: int, n_frames: int, rng=np.random, loc_std: float = 1.0, vel_norm: float = 0.5):
    """The random draws of one ChargedParticlesSim.sample_trajectory call, in its order (:228-239, :292-293): charges
    (+-1 w.p. 1/2), positions ~ N(0, (loc_std (n/5)^(1/3))^2), velocities normalised to vel_norm; then the two observation
    noise arrays are drawn (and discarded: noise_var = 0) so that the stream stays aligned from one trajectory to the
    next.  -> loc0 [3, n], vel0 [3, n], charges [n, 1]."""
    std = loc_std * (float(n_balls) / 5.0) ** (1.0 / 3.0)                       # :154-155
    charges = rng.choice(np.array([-1.0, 0.0, 1.0]), size=(n_balls, 1), p=[0.5, 0.0, 0.5])
    loc0 = rng.randn(3, n_balls) * std
    vel0 = rng.randn(3, n_balls)
    vel0 = vel0 * vel_norm / np.sqrt((vel0 ** 2).sum(axis=0)).reshape(1, -1)
    rng.randn(n_frames, 3, n_balls)
    rng.randn(n_frames, 3, n_balls)
    return loc0, vel0, charges


def _reflect(loc, vel):
    """_clamp (:192-218): elastic reflection at +-BOX, applied IN PLACE to the state the integrator continues from."""
    over = loc > BOX
    loc[over] = 2 * BOX - loc[over]
    vel[over] = -np.abs(vel[over])
    under = loc < -BOX
    loc[under] = -2 * BOX - loc[under]
    vel[under] = np.abs(vel[under])


def _charged_force(x, qq, strength, max_f):
    """x [3, n] -> F [3, n]: sum_j strength q_i q_j (x_i - x_j) / |x_i - x_j|^3 with |.|^2 expanded as
    |a|^2 + |b|^2 - 2 a.b like _l2 (:165-177), zero self term, components clipped to +-max_f (:255-273)."""
    n2 = (x ** 2).sum(axis=0)
    d2 = n2[:, None] + n2[None, :] - 2.0 * (x.T @ x)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = strength * qq / np.power(d2, 1.5)
    np.fill_diagonal(w, 0.0)
    F = (w[None, :, :] * (x[:, :, None] - x[:, None, :])).sum(axis=-1)
    return np.clip(F, -max_f, max_f)


def simulate_charged(loc0, vel0, charges, T: int, sample_freq: int, strength: float = 1.0, dt: float = DT_CHARGED):
    """-> loc, vel [T // sample_freq - 1, 3, n]: frame k is the state after (k + 1) sample_freq drift steps (the frame
    written before the loop is overwritten by the first sample, :240 vs :277-279); the stored velocity is the leapfrog
    half-step velocity."""
    assert T % sample_freq == 0
    n_save = T // sample_freq - 1
    x, v = np.array(loc0, dtype=np.float64), np.array(vel0, dtype=np.float64)
    q = np.asarray(charges, dtype=np.float64).reshape(-1, 1)
    qq = q @ q.T
    max_f = 0.1 / dt
    loc, vel = np.zeros((n_save, 3, x.shape[1])), np.zeros((n_save, 3, x.shape[1]))
    _reflect(x, v)
    v += dt * _charged_force(x, qq, strength, max_f)
    k = 0
    for i in range(1, T):
        x += dt * v
        if i % sample_freq == 0:
            loc[k], vel[k] = x, v
            k += 1
        v += dt * _charged_force(x, qq, strength, max_f)
    return loc, vel


def gravity_initial_conditions(n_balls: int, n_frames: int, rng=np.random, loc_std: float = 1.0):
    """The draws of one GravitySim.sample_trajectory call (:370-379, :401-403): masses 1 + 0.1 loc_std N(0,1), positions
    and velocities ~ N(0,1), velocities moved to the centre-of-mass frame; then three discarded noise arrays."""
    mass = np.ones((n_balls, 1)) + rng.randn(n_balls, 1) * loc_std * 0.1
    pos0 = rng.randn(n_balls, 3)
    vel0 = rng.randn(n_balls, 3)
    vel0 = vel0 - np.mean(mass * vel0, 0) / np.mean(mass)
    for _ in range(3):
        rng.randn(n_frames, n_balls, 3)
    return pos0, vel0, mass


def _gravity_acc(pos, mass, G, softening):
    """compute_acceleration (:311-333): a_i = G sum_j m_j (x_j - x_i) (|x_j - x_i|^2 + softening^2)^(-3/2)."""
    d = pos[None, :, :] - pos[:, None, :]                      # d[i, j] = x_j - x_i
    inv_r3 = ((d ** 2).sum(-1) + softening ** 2) ** (-1.5)
    return G * np.einsum("ijd,ij,j->id", d, inv_r3, mass[:, 0])


def simulate_gravity(pos0, vel0, mass, T: int, sample_freq: int, G: float = 1.0, softening: float = 0.1, dt: float = 1e-3):
    """Kick-drift-kick leapfrog (:384-399) -> pos, vel, force [T // sample_freq, n, 3]; frame k is the state BEFORE step
    k sample_freq, force = acc * mass."""
    assert T % sample_freq == 0
    n_save = T // sample_freq
    x, v = np.array(pos0, dtype=np.float64), np.array(vel0, dtype=np.float64)
    m = np.asarray(mass, dtype=np.float64).reshape(-1, 1)
    pos, vel, force = (np.zeros((n_save,) + x.shape) for _ in range(3))
    a = _gravity_acc(x, m, G, softening)
    for i in range(T):
        if i % sample_freq == 0:
            k = i // sample_freq
            pos[k], vel[k], force[k] = x, v, a * m
        v += a * dt / 2.0
        x += v * dt
        a = _gravity_acc(x, m, G, softening)
        v += a * dt / 2.0
    return pos, vel, force
