"""Tiny N-body data sets in the reference's on-disk format (TEST INFRASTRUCTURE).

File names and array layouts follow generate_dataset.py:56,134-147 of the reference:
  charged  loc / vel [S, T/freq - 1, 3, N] float64, edges [S, N, N] (= q q^T), charges [S, N, 1]
  gravity  loc / vel [S, T/freq, N, 3] float64, "edges" slot = forces [S, T/freq, N, 3], charges = masses [S, N, 1]
Initial conditions come from oracle.sim_oracle's samplers (the reference's draw order); the trajectories are integrated
by the CUDA simulator kernels when a device is given (fast; pinned to the reference by tests/test_sim.py) and by the
numpy restatement otherwise.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

from oracle import sim_oracle as S


def simulate_split(kind: str, n_balls: int, n_traj: int, length: int, sample_freq: int, seed: int, device=None):
    rng = np.random.RandomState(seed)
    if kind == "charged":
        ics = [S.charged_initial_conditions(n_balls, length // sample_freq - 1, rng=rng) for _ in range(n_traj)]
        loc0, vel0, q = (np.stack([i[k] for i in ics]) for k in range(3))
        if device is not None:
            import torch
            import no_node_comparison_b200 as nb

            t = lambda a: torch.tensor(a, dtype=torch.float64, device=device)
            loc, vel = nb.simulate_charged(t(loc0), t(vel0), t(q), length, sample_freq)
            loc, vel = loc.cpu().numpy(), vel.cpu().numpy()
        else:
            out = [S.simulate_charged(l, v, c, length, sample_freq) for l, v, c in zip(loc0, vel0, q)]
            loc, vel = np.stack([o[0] for o in out]), np.stack([o[1] for o in out])
        edges = q @ q.transpose(0, 2, 1)
        return loc, vel, edges, q
    if kind == "gravity":
        ics = [S.gravity_initial_conditions(n_balls, length // sample_freq, rng=rng) for _ in range(n_traj)]
        pos0, vel0, m = (np.stack([i[k] for i in ics]) for k in range(3))
        if device is not None:
            import torch
            import no_node_comparison_b200 as nb

            t = lambda a: torch.tensor(a, dtype=torch.float64, device=device)
            pos, vel, force = (o.cpu().numpy() for o in nb.simulate_gravity(t(pos0), t(vel0), t(m), length, sample_freq))
        else:
            out = [S.simulate_gravity(p, v, mm, length, sample_freq) for p, v, mm in zip(pos0, vel0, m)]
            pos, vel, force = (np.stack([o[k] for o in out]) for k in range(3))
        return pos, vel, force, m
    raise ValueError(kind)


def write_dataset(out_dir, kind: str, n_balls: int, splits: dict, length: int, sample_freq: int = 100, seed: int = 43,
                  device=None) -> Path:
    """splits: {'train': n, 'valid': n, 'test': n} -> the 4 x len(splits) .npy files under `out_dir`."""
    out = Path(out_dir)
    os.makedirs(out, exist_ok=True)
    suffix = f"_{kind}{n_balls}_initvel1small"
    for k, (part, n) in enumerate(splits.items()):
        loc, vel, edges, charges = simulate_split(kind, n_balls, n, length, sample_freq, seed + 1000 * k, device)
        np.save(out / f"loc_{part}{suffix}.npy", loc)
        np.save(out / f"vel_{part}{suffix}.npy", vel)
        np.save(out / f"edges_{part}{suffix}.npy", edges)
        np.save(out / f"charges_{part}{suffix}.npy", charges)
    return out
