#!/usr/bin/env python
"""Throughput of the device trajectory simulators next to the oracle's numpy loop (the reference's algorithm) on a few
trajectories: python tools/sim_throughput.py  (GPU box)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import no_node_comparison_b200 as nb  # noqa: E402
from oracle import sim_oracle as S  # noqa: E402

dev = torch.device("cuda:0")
for N, B, T in ((20, 512, 5000), (100, 512, 5000)):
    rng = np.random.RandomState(1)
    ics = [S.charged_initial_conditions(N, T // 100 - 1, rng) for _ in range(B)]
    l0 = torch.tensor(np.stack([i[0] for i in ics]), device=dev)
    v0 = torch.tensor(np.stack([i[1] for i in ics]), device=dev)
    q = torch.tensor(np.stack([i[2] for i in ics]), device=dev)
    nb.simulate_charged(l0, v0, q, 200, 100)
    torch.cuda.synchronize()
    t0 = time.time()
    loc, vel = nb.simulate_charged(l0, v0, q, T, 100)
    torch.cuda.synchronize()
    dt = time.time() - t0
    t1 = time.time()
    ref_l, _ = S.simulate_charged(*ics[0], T, 100)
    dc = time.time() - t1
    err = np.abs(loc[0].cpu().numpy() - ref_l).max()
    print(f"charged N={N}: {B} trajectories x {T} steps in {dt:.3f} s = {B / dt:.0f} trajectories/s on the GPU; numpy loop "
          f"{1 / dc:.2f} trajectories/s on one host core; max |diff| of trajectory 0 = {err:.2e}")
