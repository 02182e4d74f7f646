"""Device-resident featurisation, conserved-energy evaluation and autoregressive rollout.

The reference's callers rebuild the model inputs with ~10 small torch ops per call (prepare_inputs,
EGNO/main_simulation_simple_no.py:311-339; SEGNO/train_nbody.py:119-123, :228-233) and evaluate the conserved energy in
numpy on the host once per emitted frame (utils.py:197-219 -> a device->host synchronisation per frame).  Here both are
single kernels behind the C ABI (nb_nbody_features, nb_nbody_energy), so `egno_rollout` / `segno_rollout` — the
counterparts of the two `rollout_fn`s (main_simulation_simple_no.py:342-384, train_nbody.py:200-236) for the
single-input case — enqueue the whole long-horizon rollout without ever synchronising; the caller reads the results once.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from ._lib import check, load_library
from .functional import _ptr, _require_cuda_f32, _stream_ptr

_KIND = {"charged": 0, "gravity": 1}


def prepare_inputs(loc: torch.Tensor, vel: torch.Tensor, charges: torch.Tensor, n_nodes: int, with_charge: bool = True,
                   edge_attr_o: Optional[torch.Tensor] = None, want_mean: bool = True):
    """loc, vel [B*N, 3], charges [B*N(,1)] -> (nodes [B*N, 1 + with_charge], loc_mean [B*N, 3] | None,
    edge_attr [B*N*(N-1), 2]) for the canonical fully connected edge order."""
    lib = load_library()
    bn = loc.shape[0]
    if bn % n_nodes:
        raise ValueError(f"{bn} nodes are not a multiple of n_nodes={n_nodes}")
    B, N = bn // n_nodes, n_nodes
    loc = _require_cuda_f32("loc", loc, (bn, 3))
    vel = _require_cuda_f32("vel", vel, (bn, 3))
    q = _require_cuda_f32("charges", charges.reshape(-1), (bn,))
    E = bn * (N - 1)
    eo = None if edge_attr_o is None else _require_cuda_f32("edge_attr_o", edge_attr_o.reshape(-1), (E,))
    dev = loc.device
    nodes = torch.empty((bn, 2 if with_charge else 1), device=dev, dtype=torch.float32)
    mean = torch.empty((bn, 3), device=dev, dtype=torch.float32) if want_mean else None
    ea = torch.empty((E, 2), device=dev, dtype=torch.float32)
    check(lib.nb_nbody_features(B, N, int(with_charge), _ptr(loc), _ptr(vel), _ptr(q), _ptr(eo), _ptr(nodes), _ptr(mean),
                                _ptr(ea), _stream_ptr(dev)), "nb_nbody_features")
    return nodes, mean, ea


def conserved_energy(dataset: str, loc: torch.Tensor, vel: torch.Tensor, charges: torch.Tensor, n_nodes: int,
                     G: float = 1.0) -> torch.Tensor:
    """loc, vel [F, B*N, 3] (or [B*N, 3]), charges / masses [B*N(,1)] -> energy [F, B] on the device
    (utils.py:126-144 charged, :175-195 gravity)."""
    if dataset not in _KIND:
        raise ValueError(f"unknown dataset {dataset!r} (charged | gravity)")
    lib = load_library()
    if loc.dim() == 2:
        loc, vel = loc[None], vel[None]
    F, bn = loc.shape[0], loc.shape[1]
    if bn % n_nodes:
        raise ValueError(f"{bn} nodes are not a multiple of n_nodes={n_nodes}")
    B = bn // n_nodes
    loc = _require_cuda_f32("loc", loc, (F, bn, 3))
    vel = _require_cuda_f32("vel", vel, (F, bn, 3))
    q = _require_cuda_f32("charges", charges.reshape(-1), (bn,))
    out = torch.empty((F, B), device=loc.device, dtype=torch.float32)
    check(lib.nb_nbody_energy(_KIND[dataset], F, B, n_nodes, ctypes.c_float(G), _ptr(loc), _ptr(vel), _ptr(q), _ptr(out),
                              _stream_ptr(loc.device)), "nb_nbody_energy")
    return out


@torch.no_grad()
def egno_rollout(model, loc: torch.Tensor, vel: torch.Tensor, charges: torch.Tensor, edges, n_nodes: int, traj_len: int,
                 dataset: Optional[str] = "charged", edge_attr_o: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """EGNO long-horizon rollout, num_inputs == 1 (main_simulation_simple_no.py:342-384): `traj_len` model calls of
    T = model.num_timesteps frames each, the last frame of a call seeding the next.
    -> loc_preds [traj_len*T, B*N, 3], energies [traj_len, B] (last frame of each call), energies_allsteps
    [traj_len*T, B]; the energies are None when `dataset` is None.  Nothing synchronises with the host."""
    T = model.num_timesteps
    bn = loc.shape[0]
    B = bn // n_nodes
    t_out = torch.arange(1, T + 1, device=loc.device)[None].repeat(B, 1)
    preds = torch.empty((traj_len, T, bn, 3), device=loc.device, dtype=torch.float32)
    vels = torch.empty((traj_len, T, bn, 3), device=loc.device, dtype=torch.float32)
    for i in range(traj_len):
        nodes, mean, ea = prepare_inputs(loc, vel, charges, n_nodes, True, edge_attr_o)
        x, v, _ = model(loc, nodes, edges, ea, v=vel, loc_mean=mean, timesteps_out=t_out)
        preds[i], vels[i] = x.view(T, bn, 3), v.view(T, bn, 3)
        loc, vel = preds[i, T - 1], vels[i, T - 1]
    e_all = e_last = None
    if dataset is not None:
        e_all = conserved_energy(dataset, preds.view(traj_len * T, bn, 3), vels.view(traj_len * T, bn, 3), charges, n_nodes)
        e_last = e_all.view(traj_len, T, B)[:, T - 1]
    return preds.view(traj_len * T, bn, 3), e_last, e_all


@torch.no_grad()
def segno_rollout(model, loc: torch.Tensor, vel: torch.Tensor, charges: torch.Tensor, edges, n_nodes: int, traj_len: int,
                  num_steps: int = 10, dataset: Optional[str] = "gravity"
                  ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """SEGNO long-horizon rollout, num_prev == 1 (train_nbody.py:200-236): `traj_len` calls of `num_steps` integration
    sub-steps, one frame out per call.  -> loc_preds [traj_len, B*N, 3], energies [traj_len, B] | None."""
    bn = loc.shape[0]
    preds = torch.empty((traj_len, bn, 3), device=loc.device, dtype=torch.float32)
    vels = torch.empty((traj_len, bn, 3), device=loc.device, dtype=torch.float32)
    for i in range(traj_len):
        his, _, ea = prepare_inputs(loc, vel, charges, n_nodes, False, None, want_mean=False)
        x, _, v = model(his, loc, edges, vel, ea, T=num_steps)
        preds[i], vels[i] = x, v
        loc, vel = preds[i], vels[i]
    en = conserved_energy(dataset, preds, vels, charges, n_nodes) if dataset is not None else None
    return preds, en
