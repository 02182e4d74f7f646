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
                 weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat_state = {}

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
            st = self._flat_state.get(gi)
            total = sum(p.numel() for p in ps)
            if st is None:
                st = dict(m=torch.zeros(total, device=dev), v=torch.zeros(total, device=dev),
                          step=torch.zeros(1, device=dev), offs={})
                o = 0
                for p in ps:
                    st["offs"][id(p)] = o
                    o += p.numel()
                self._flat_state[gi] = st
            b1, b2 = group["betas"]
            tick = 1
            for first, n in self._runs(ps):
                o = st["offs"][id(first)]
                check(lib.nb_adam_step(n, _ptr(first.data), _ptr(first.grad), _ptr(st["m"][o:o + n]), _ptr(st["v"][o:o + n]),
                                       _ptr(st["step"]), tick, ctypes.c_double(group["lr"]), ctypes.c_double(b1),
                                       ctypes.c_double(b2), ctypes.c_double(group["eps"]), ctypes.c_double(group["weight_decay"]),
                                       _stream_ptr(dev)), "nb_adam_step")
                tick = 0
        return loss
