"""FlatAdam — torch.optim.Adam semantics as ONE fused kernel over the model's flat parameter / gradient buffers
(the reference builds `torch.optim.Adam(model.parameters(), lr, weight_decay)` at main.py:150).

The drop-in modules keep their parameters as views of one flat fp32 buffer and their backward writes one flat gradient
buffer (functional._ParamPack), so an optimizer step is a single launch (plus a one-thread step-counter tick) instead of
one multi-tensor kernel sequence; the step counter lives on the device, so the step is CUDA-graph capturable.
Parameters whose `.grad` is None are skipped, as in torch.optim.Adam.  If the gradients are not the contiguous views this
relies on (e.g. after gradient accumulation into separately allocated tensors) the update runs per tensor instead.
"""
from __future__ import annotations

import ctypes
from typing import List, Tuple

import torch

from ._lib import check, load_library
from .functional import _ptr, _stream_ptr


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, grad_scale: float = 1.0, peer_bucket=None):
        """`grad_scale`: the gradients are multiplied by it inside the update kernel (data parallel: 1 / world size when
        the flat gradient bucket holds the SUM over ranks).  `peer_bucket` (model.peer_bucket after
        enable_data_parallel(peer_memory=True)): the gradient all-reduce happens INSIDE the update kernel — it sums every
        rank's flat gradient buffer through peer memory, in rank order, times grad_scale / world."""
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_scale=grad_scale))
        self.peer_bucket = peer_bucket
        self._flat = {}     # group index -> dict(m, v, step, offs): flat moment buffers; self.state[p] holds views of them

    @staticmethod
    def _runs(ps: List[torch.nn.Parameter]):
        """Maximal runs of parameters whose data AND grads are back to back in memory -> [(first_param, n_elements)]."""
        runs, cur, n = [], None, 0
        for p in ps:
            if p.grad is None:
                if cur is not None:
                    runs.append((cur, n))
                cur, n = None, 0
                continue
            ok = (cur is not None and p.data_ptr() == cur.data_ptr() + 4 * n and p.grad.data_ptr() == cur.grad.data_ptr() + 4 * n
                  and p.is_contiguous() and p.grad.is_contiguous())
            if ok:
                n += p.numel()
            else:
                if cur is not None:
                    runs.append((cur, n))
                cur, n = p, p.numel()
        if cur is not None:
            runs.append((cur, n))
        return runs

    def _group_state(self, gi: int, ps: List[torch.nn.Parameter]):
        """Flat exp_avg / exp_avg_sq / step of a parameter group.  `self.state[p]` holds views of the flat buffers under
        torch.optim.Adam's keys, so `state_dict()` / `load_state_dict()` round-trip them; whenever the per-parameter
        state no longer aliases the flat buffers (after load_state_dict, add_param_group or a re-ordered group) the flat
        buffers are rebuilt from it."""
        st = self._flat.get(gi)
        dev = ps[0].device
        total = sum(p.numel() for p in ps)

        def aliased():
            if st is None or st["m"].numel() != total or st["m"].device != dev or len(st["offs"]) != len(ps):
                return False
            o = 0
            for p in ps:
                s = self.state.get(p)
                if st["offs"].get(id(p)) != o:
                    return False
                if s:   # parameters that never received a gradient have no state yet
                    if (s["exp_avg"].data_ptr() != st["m"].data_ptr() + 4 * o or s["exp_avg_sq"].data_ptr() != st["v"].data_ptr() + 4 * o
                            or s["step"].data_ptr() != st["step"].data_ptr()):
                        return False
                o += p.numel()
            return True

        if aliased():
            return st
        new = dict(m=torch.zeros(total, device=dev), v=torch.zeros(total, device=dev), step=torch.zeros(1, device=dev), offs={})
        o, step_src = 0, None
        for p in ps:
            new["offs"][id(p)] = o
            s = self.state.get(p)
            if s:   # carry over what a checkpoint (or an earlier layout) holds
                new["m"][o:o + p.numel()].copy_(s["exp_avg"].reshape(-1))
                new["v"][o:o + p.numel()].copy_(s["exp_avg_sq"].reshape(-1))
                if step_src is None:
                    step_src = s["step"]
            o += p.numel()
        if step_src is not None:
            new["step"].copy_(torch.as_tensor(step_src, dtype=torch.float32).reshape(1))
        self._flat[gi] = new
        return new

    def _publish(self, st, p):
        o = st["offs"][id(p)]
        self.state[p] = dict(step=st["step"], exp_avg=st["m"][o:o + p.numel()].view(p.shape),
                             exp_avg_sq=st["v"][o:o + p.numel()].view(p.shape))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = load_library()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"]]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32:
                    raise ValueError("FlatAdam needs float32 CUDA parameters")
                if p.grad is not None and (p.grad.dtype != torch.float32 or p.grad.is_sparse):
                    raise ValueError("FlatAdam needs dense float32 gradients")
            dev = ps[0].device
            st = self._group_state(gi, ps)
            for p in ps:
                if p.grad is not None and not self.state.get(p):
                    self._publish(st, p)
            b1, b2 = group["betas"]
            tick = 1
            pb = self.peer_bucket
            if pb is not None:
                if len(self.param_groups) != 1 or sum(p.numel() for p in ps) != pb.numel:
                    raise ValueError("peer_bucket needs ONE parameter group holding exactly the model's parameters")
                pb.barrier()        # every rank's backward has written its bucket
            for first, n in self._runs(ps):
                o = st["offs"][id(first)]
                hyper = (ctypes.c_double(group["lr"]), ctypes.c_double(b1), ctypes.c_double(b2), ctypes.c_double(group["eps"]),
                         ctypes.c_double(group["weight_decay"]))
                if pb is None:
                    check(lib.nb_adam_step(n, _ptr(first.data), _ptr(first.grad), _ptr(st["m"][o:o + n]), _ptr(st["v"][o:o + n]),
                                           _ptr(st["step"]), tick, *hyper, ctypes.c_double(group.get("grad_scale", 1.0)),
                                           _stream_ptr(dev)), "nb_adam_step")
                else:
                    if first.grad.data_ptr() != pb.buf.data_ptr() + 4 * o:
                        raise RuntimeError("peer-memory data parallel: .grad does not alias the model's gradient bucket "
                                           "(use zero_grad(set_to_none=True) and one backward per step)")
                    peers = (ctypes.c_void_p * pb.world)(*[ptr + 4 * o for ptr in pb.ptrs])
                    check(lib.nb_adam_step_peers(n, _ptr(first.data), peers, pb.world, _ptr(st["m"][o:o + n]), _ptr(st["v"][o:o + n]),
                                                 _ptr(st["step"]), tick, *hyper,
                                                 ctypes.c_double(group.get("grad_scale", 1.0) / pb.world), _stream_ptr(dev)),
                          "nb_adam_step_peers")
                tick = 0
            if pb is not None:
                pb.barrier()        # nobody overwrites a bucket (next backward) before every rank has read it
        return loss
