"""Synthetic N-body batches for throughput runs and large-size property tests (host-side harness).

Samples the *initial-condition distribution* of the reference simulators (no time integration; runtime of the
model is value-independent) and builds model inputs the way the reference callers do:

  charged (synthetic_sim.py:149-232): loc ~ N(0, loc_std^2), loc_std = (N/5)^(1/3); |v| = 0.5; q = +-1 w.p. 1/2
  gravity (synthetic_sim.py:367-382): loc, vel ~ N(0, 1), centre-of-mass velocity removed; m = 1 + 0.1 N(0,1)

  EGNO features   = prepare_inputs           (EGNO/main_simulation_simple_no.py:329-338)
  SEGNO features  = run_epoch featurisation  (SEGNO/train_nbody.py:93,119-123)
"""
from __future__ import annotations

from typing import Dict

import torch


def canonical_edges(batch: int, n_nodes: int, device="cpu"):
    """`for i: for j != i` per graph, graphs offset by N*b (EGNO/simulation/dataset_simple.py:64-71, :101-111)."""
    i = torch.arange(n_nodes, device=device).repeat_interleave(n_nodes)
    j = torch.arange(n_nodes, device=device).repeat(n_nodes)
    keep = i != j
    i, j = i[keep], j[keep]
    off = (torch.arange(batch, device=device) * n_nodes).repeat_interleave(i.numel())
    return i.repeat(batch) + off, j.repeat(batch) + off


def sample_state(kind: str, batch: int, n_nodes: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """-> loc, vel [B,N,3] float32, charges [B,N,1] (charges or masses), on the host."""
    g = torch.Generator().manual_seed(seed)
    if kind == "charged":
        loc_std = (n_nodes / 5.0) ** (1.0 / 3.0)
        loc = torch.randn(batch, n_nodes, 3, generator=g) * loc_std
        vel = torch.randn(batch, n_nodes, 3, generator=g)
        vel = vel * 0.5 / vel.norm(dim=-1, keepdim=True)
        q = (torch.randint(0, 2, (batch, n_nodes, 1), generator=g).float() * 2 - 1)
    elif kind == "gravity":
        q = 1.0 + 0.1 * torch.randn(batch, n_nodes, 1, generator=g)
        loc = torch.randn(batch, n_nodes, 3, generator=g)
        vel = torch.randn(batch, n_nodes, 3, generator=g)
        vel = vel - (q * vel).mean(1, keepdim=True) / q.mean(1, keepdim=True)
    else:
        raise ValueError(kind)
    return dict(loc=loc, vel=vel, charges=q)


def egno_features(loc, vel, charges, row, col):
    """[B,N,3] states -> (x[BN,3], nodes[BN,2], edge_attr[E,2], v[BN,3], loc_mean[BN,3])."""
    B, N, _ = loc.shape
    loc_mean = loc.mean(dim=1, keepdim=True).expand(B, N, 3).reshape(-1, 3).contiguous()
    x = loc.reshape(-1, 3)
    v = vel.reshape(-1, 3)
    q = charges.reshape(-1, 1)
    nodes = torch.cat([v.norm(dim=1, keepdim=True), q], dim=1)
    d = x[row] - x[col]
    edge_attr = torch.cat([q[row] * q[col], (d * d).sum(1, keepdim=True)], dim=1)
    return x.contiguous(), nodes.contiguous(), edge_attr.contiguous(), v.contiguous(), loc_mean


def segno_features(loc, vel, charges, row, col):
    """[B,N,3] states -> (his[BN,1], x[BN,3], v[BN,3], edge_attr[E,2])."""
    x = loc.reshape(-1, 3)
    v = vel.reshape(-1, 3)
    q = charges.reshape(-1, 1)
    d = x[row] - x[col]
    edge_attr = torch.cat([q[row] * q[col], (d * d).sum(1, keepdim=True)], dim=1)
    return v.norm(dim=1, keepdim=True).contiguous(), x.contiguous(), v.contiguous(), edge_attr.contiguous()
