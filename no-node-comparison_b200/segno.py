"""Drop-in `SEGNO` module: constructor kwargs, forward signature, parameter names and shapes of the reference
class (SEGNO/models/model.py:6-102 with SEGNO_GCL, SEGNO/models/models/gcl.py:26-119).

Served semantics: `forward` returns the integrated state `forward_step(embedding(his), ...)` in the
reference's order (x, h, v).  The literal `forward` at reference HEAD (model.py:53-92) returns its *inputs*
and cannot be trained; the dead first definition (model.py:28-51), `rollout_fn` and the
training loop all assume the integrated state, which is what is computed here.
"""
from __future__ import annotations

import torch
from torch import nn

from .functional import SegnoFunction, _EdgeCache, _ParamPack, _require_cuda_f32

_HIDDEN = 64


class _GCLHolder(nn.Module):
    """Parameters of SEGNO_GCL created in the reference's RNG order (gcl.py:39-67)."""

    def __init__(self, hidden: int, edges_in_d: int, act_fn: nn.Module):
        super().__init__()
        self.edge_mlp = nn.Sequential(nn.Linear(2 * hidden + 1 + edges_in_d, hidden), act_fn,
                                      nn.Linear(hidden, hidden), act_fn)
        self.node_mlp = nn.Sequential(nn.Linear(hidden + hidden, hidden), act_fn, nn.Linear(hidden, hidden))
        last = nn.Linear(hidden, 1, bias=True)                      # created BEFORE coord_mlp.0 (gcl.py:50-54)
        torch.nn.init.xavier_uniform_(last.weight, gain=0.001)
        self.coord_mlp = nn.Sequential(nn.Linear(hidden, hidden), act_fn, last)
        self.coord_mlp_vel = nn.Sequential(nn.Linear(hidden, hidden), act_fn, nn.Linear(hidden, 1))  # inert (gcl.py:64-67)
        self.n_layers = 8


class SEGNO(nn.Module):
    def __init__(self, in_node_nf, in_edge_nf, hidden_nf, device='cpu', act_fn=nn.SiLU(), n_layers=4, coords_weight=1.0,
                 recurrent=False, norm_diff=False, tanh=False, invariant=True, norm_vel=True, varDT=False,
                 multiple_agg=None):
        super().__init__()
        if hidden_nf != _HIDDEN:
            raise ValueError(f"kernels are specialised for hidden_nf={_HIDDEN} (model_confs.yaml:23), got {hidden_nf}")
        if not isinstance(act_fn, nn.SiLU):
            raise ValueError("only the SiLU activation (reference default) is implemented")
        if tanh:
            raise ValueError("tanh=True (gcl.py:57-59) is not implemented (model_confs.yaml:27 uses False)")
        if multiple_agg is not None:
            raise NotImplementedError("multi-input SEGNO (multiple_agg, model.py:70-90,105-139) is not implemented yet")
        self.hidden_nf = hidden_nf
        self.varDT = varDT
        self.multiple_agg = multiple_agg
        self.device = device
        self.n_layers = n_layers
        self.in_node_nf = in_node_nf
        self.in_edge_nf = in_edge_nf
        self.recurrent = recurrent
        self.coords_weight = coords_weight
        self.embedding = nn.Linear(in_node_nf, hidden_nf)
        self.invariant = invariant
        self.norm_vel = norm_vel
        self.module = _GCLHolder(hidden_nf, in_edge_nf, act_fn)
        self.module.n_layers = n_layers
        self.to(self.device)
        self._pack = _ParamPack(self)
        self._edges = _EdgeCache()
        self.process_group = None

    def enable_data_parallel(self, group=None):
        import torch.distributed as dist

        self.process_group = group if group is not None else dist.group.WORLD
        return self

    def forward(self, his, x, edges, v, edge_attr, T=10, in_steps=None):
        if x.dim() == 3:
            raise NotImplementedError("multi-input SEGNO (x of shape [BN, n_inputs, 3]) is not implemented yet")
        if edge_attr.requires_grad or his.requires_grad:
            raise ValueError("gradients w.r.t. his / edge_attr are not implemented (the reference callers detach them)")
        T = int(T)
        n0 = x.shape[0]
        dev = x.device
        x = _require_cuda_f32("x", x, (n0, 3))
        v = _require_cuda_f32("v", v, (n0, 3))
        his = _require_cuda_f32("his", his, (n0, self.in_node_nf))
        E = edge_attr.shape[0]
        # E = B*N*(N-1) and n0 = B*N  ->  N - 1 = E / n0
        if n0 == 0 or E % n0 != 0:
            raise ValueError(f"{E} edges / {n0} nodes is not a batch of fully connected graphs")
        N = E // n0 + 1
        if n0 % N != 0:
            raise ValueError(f"{n0} nodes do not divide into graphs of N={N}")
        B = n0 // N
        edge_attr = _require_cuda_f32("edge_attr", edge_attr, (E, self.in_edge_nf))
        self._edges.validate(edges, B, N, dev)
        # forward_step mutates these on every call (model.py:96-97)
        self.module.n_layers = T
        self.n_layers = T
        cfg = (B, N, T, self.in_node_nf, self.in_edge_nf, 1 if self.recurrent else 0, float(self.coords_weight))
        from ._lib import load_library, check
        from . import _cabi
        import ctypes
        expected = load_library().nb_segno_param_count(ctypes.byref(_cabi.NbSegnoConfig(*cfg)))
        if expected < 0:
            check(-1, "SEGNO configuration")
        flat, params = self._pack.flat_params(expected, dev)
        return SegnoFunction.apply(cfg, self.process_group, flat, his, x, v, edge_attr, *params)

    def forward_step(self, h, x, edges, v, edge_attr, T=10):
        raise NotImplementedError("call forward(his, ...): embedding and the T integration steps run as one fused call")
