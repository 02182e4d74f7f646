"""Fused trajectory MSE: the loss of the reference's training loops and its gradient in two launches.

EGNO/main_simulation_simple_no.py:268-276 reshapes the frame-major prediction to [B, N, T, 3], takes
`MSELoss(reduction='none')(pred, target).mean((0, 1, 3))` -> losses[T] and then `losses.mean()` (or `losses[0]`);
SEGNO/train_nbody.py:163-165 is the same with one frame.  Autograd turns that into ~10 element-wise / reduction
launches over 600 KB tensors; here one kernel reads prediction and target once, writes d loss / d pred and per-CTA
partial sums, and a second one-CTA kernel reduces them in a fixed order (deterministic, CUDA-graph capturable).
"""
from __future__ import annotations

import torch

from ._lib import check, load_library
from .functional import _ptr, _require_cuda_f32, _stream_ptr


class _TrajectoryMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, T, layout, only_first):
        lib = load_library()
        rows = pred.shape[0] // T
        dev = pred.device
        losses = torch.empty(T, device=dev, dtype=torch.float32)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(pred) if need_grad else None
        ws = torch.empty(lib.nb_traj_mse_workspace_floats(T), device=dev, dtype=torch.float32)
        check(lib.nb_traj_mse(T, rows, layout, int(only_first), _ptr(pred), _ptr(target), _ptr(losses), _ptr(loss), _ptr(grad),
                              _ptr(ws), _stream_ptr(dev)), "nb_traj_mse")
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(losses)
        return loss, losses

    @staticmethod
    def backward(ctx, g_loss, _g_losses):
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None, None, None


def trajectory_mse(pred: torch.Tensor, target: torch.Tensor, num_timesteps: int = 1, only_first: bool = False):
    """pred [T*B*N, 3] (the models' frame-major output; T = 1 for SEGNO), target [T*B*N, 3] in the same order or
    [B*N, T, 3] / [B, N, T, 3] as the reference's loader holds it -> (loss, losses[T]).  loss is differentiable with
    respect to pred; the target takes no gradient."""
    T = int(num_timesteps)
    if pred.dim() != 2 or pred.shape[1] != 3 or pred.shape[0] % T:
        raise ValueError(f"pred must be [T*B*N, 3] with T={T}, got {tuple(pred.shape)}")
    rows = pred.shape[0] // T
    if target.requires_grad:
        raise ValueError("gradients w.r.t. the target are not implemented")
    if target.dim() == 2:
        layout, tshape = 0, (T * rows, 3)
    elif target.dim() in (3, 4) and target.shape[-2] == T:
        layout, tshape = 1, tuple(target.shape)
        if target.numel() != T * rows * 3:
            raise ValueError(f"target {tuple(target.shape)} does not match pred {tuple(pred.shape)}")
    else:
        raise ValueError(f"target must be [T*B*N, 3], [B*N, T, 3] or [B, N, T, 3], got {tuple(target.shape)}")
    pred_c = _require_cuda_f32("pred", pred, (T * rows, 3))
    target_c = _require_cuda_f32("target", target, tshape)
    return _TrajectoryMSE.apply(pred_c, target_c, T, layout, bool(only_first))
