"""The library's second stream (weight-gradient reductions underneath the next layer's kernels) against the
synchronous flush on the caller's stream.  The switch NB_B200_WGRAD_DEFER is read once per process, so each arm runs in
its own interpreter on the same seeded case and writes outputs and gradients to a file.

EGNO: every reduction keeps its order and its operands, so the two arms must agree BITWISE (graph replay against eager is
test_cuda_graph_training_step_matches_eager).  SEGNO: the deferred arm cuts the row range of a reduction into chunks of sub-steps (deterministic, but a different
summation tree): 1e-5 of the tensor's scale, and bitwise between two runs of the same arm."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
import no_node_comparison_b200 as nb
from no_node_comparison_b200 import synth
out, model = sys.argv[2], sys.argv[3]
dev = torch.device("cuda:0")
res = {}
if model == "egno":
    B, N, T = 24, 20, 10
    s = synth.sample_state("charged", B, N, seed=3)
    row, col = synth.canonical_edges(B, N, dev)
    x, nodes, ea, v, lm = synth.egno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), row, col)
    torch.manual_seed(5)
    m = nb.EGNO(n_layers=4, in_node_nf=2, in_edge_nf=2, hidden_nf=64, device=dev, with_v=True, num_modes=2,
                num_timesteps=T, time_emb_dim=32)
    t_out = torch.arange(1, T + 1, device=dev)[None].repeat(B, 1)
    target = torch.randn(B * N, T, 3, device=dev)

    def step():
        m.zero_grad(set_to_none=True)
        xo, vo, ho = m(x, nodes, [row, col], ea, v=v, loc_mean=lm, timesteps_out=t_out)
        loss = nb.trajectory_mse(xo, target.reshape(B, N, T, 3), T)[0] + 1e-3 * ho.square().mean() + 1e-2 * vo.square().mean()
        loss.backward()
        return loss
    for rep in range(2):
        loss = step()
        res[f"loss{rep}"] = loss.detach().cpu().numpy()
        for k, p in m.named_parameters():
            res[f"g{rep}:{k}"] = p.grad.detach().cpu().numpy()
else:
    B, N, T = 40, 20, 10
    s = synth.sample_state("gravity", B, N, seed=4)
    row, col = synth.canonical_edges(B, N, dev)
    his, x, v, ea = synth.segno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), row, col)
    torch.manual_seed(6)
    m = nb.SEGNO(in_node_nf=1, in_edge_nf=2, hidden_nf=64, device=dev, n_layers=8, recurrent=True)
    for rep in range(2):
        m.zero_grad(set_to_none=True)
        xg = x.clone().requires_grad_(True)
        xo, ho, vo = m(his, xg, [row, col], v, ea, T=T)
        (xo.square().sum() + 0.1 * ho.sum() + vo.square().sum()).backward()
        res[f"gx{rep}"] = xg.grad.cpu().numpy()
        for k, p in m.named_parameters():
            if p.grad is not None:
                res[f"g{rep}:{k}"] = p.grad.detach().cpu().numpy()
torch.cuda.synchronize()
np.savez(out, **res)
"""


def _arm(tmp_path, model, defer):
    out = str(tmp_path / f"{model}_{defer}.npz")
    env = dict(os.environ, NB_B200_WGRAD_DEFER=str(defer))
    r = subprocess.run([sys.executable, "-c", _CHILD, ROOT, out, model], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    z = np.load(out)
    return {k: z[k] for k in z.files}


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["egno", "segno"])
def test_deferred_reductions_match_the_synchronous_flush(tmp_path, model):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no fallback exists)"
    a, b = _arm(tmp_path, model, 1), _arm(tmp_path, model, 0)
    assert a.keys() == b.keys() and len(a) > 10
    for k in a:
        if k.startswith("g1:") or k in ("gx1", "loss1"):   # second call of the same arm: bitwise the first
            assert np.array_equal(a[k], a[k.replace("1", "0", 1)]), k
        if model == "egno":
            assert np.array_equal(a[k], b[k]), k
        else:
            sc = max(float(np.abs(b[k]).max()), 1e-30)
            assert float(np.abs(a[k] - b[k]).max()) <= 1e-5 * sc, (k, float(np.abs(a[k] - b[k]).max()) / sc)
