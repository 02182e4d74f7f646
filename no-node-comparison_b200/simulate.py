"""The reference's trajectory simulators on the device (data generation).

`ChargedParticlesSim.sample_trajectory` (synthetic_sim.py:220-296) and `GravitySim.sample_trajectory` (:360-405) advance
one trajectory at a time in numpy (~1 s per 100-body trajectory, 20 000 sequential steps); `simulate_charged` /
`simulate_gravity` integrate all trajectories of a data set together in float64, one CTA per trajectory, and return the
frames the reference stores, in its layouts.  Random draws (initial conditions, observation noise) belong to the caller
and stay on the host in the reference's order, so a data set regenerated here uses the same stream as
generate_dataset.py (tests/golden/make_sim_golden.py shows the draw order).
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import check, load_library
from .functional import _ptr, _stream_ptr


def _f64(name, t, shape):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (this implementation has no CPU path)")
    if t.dtype != torch.float64:
        raise ValueError(f"{name} must be float64 like the reference simulator, got {t.dtype}")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t.contiguous()


def simulate_charged(loc0, vel0, charges, T: int, sample_freq: int, interaction_strength: float = 1.0, dt: float = 1e-3,
                     box_size: float = 5.0):
    """loc0, vel0 [B, 3, N], charges [B, N(, 1)] -> loc, vel [B, T // sample_freq - 1, 3, N] (the reference's frames)."""
    if T % sample_freq or T < 2 * sample_freq:
        raise ValueError("T must be a multiple of sample_freq and hold at least one stored frame")
    B, _, N = loc0.shape
    loc0 = _f64("loc0", loc0, (B, 3, N))
    vel0 = _f64("vel0", vel0, (B, 3, N))
    q = _f64("charges", charges.reshape(B, N), (B, N))
    ns = T // sample_freq - 1
    loc = torch.empty((B, ns, 3, N), device=loc0.device, dtype=torch.float64)
    vel = torch.empty_like(loc)
    c = ctypes.c_double
    check(load_library().nb_sim_charged(B, N, T, sample_freq, c(dt), c(interaction_strength), c(0.1 / dt), c(box_size), _ptr(loc0),
                                        _ptr(vel0), _ptr(q), _ptr(loc), _ptr(vel), _stream_ptr(loc0.device)), "nb_sim_charged")
    return loc, vel


def simulate_gravity(pos0, vel0, mass, T: int, sample_freq: int, G: float = 1.0, softening: float = 0.1, dt: float = 1e-3):
    """pos0, vel0 [B, N, 3], mass [B, N(, 1)] -> pos, vel, force [B, T // sample_freq, N, 3]."""
    if T % sample_freq or T < sample_freq:
        raise ValueError("T must be a positive multiple of sample_freq")
    B, N, _ = pos0.shape
    pos0 = _f64("pos0", pos0, (B, N, 3))
    vel0 = _f64("vel0", vel0, (B, N, 3))
    m = _f64("mass", mass.reshape(B, N), (B, N))
    ns = T // sample_freq
    pos = torch.empty((B, ns, N, 3), device=pos0.device, dtype=torch.float64)
    vel, force = torch.empty_like(pos), torch.empty_like(pos)
    c = ctypes.c_double
    check(load_library().nb_sim_gravity(B, N, T, sample_freq, c(dt), c(G), c(softening), _ptr(pos0), _ptr(vel0), _ptr(m),
                                        _ptr(pos), _ptr(vel), _ptr(force), _stream_ptr(pos0.device)), "nb_sim_gravity")
    return pos, vel, force
