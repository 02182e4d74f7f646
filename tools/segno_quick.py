#!/usr/bin/env python
"""SEGNO timing at the configs[3] shape (gravity, N=20, 10 sub-steps, B=256): graph-replayed training step, inference,
and the per-kernel means of an eager profiled pass.  One JSON line; for same-box A/B of library variants
(NB_B200_LIBRARY / NB_B200_* switches) without the rest of bench.py.

    python tools/segno_quick.py [steps]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import no_node_comparison_b200 as nb
from no_node_comparison_b200 import synth

dev = torch.device("cuda:0")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B, N, T = 256, 20, 10
row, col = synth.canonical_edges(B, N, dev)
batches = []
for i in range(4):
    s = synth.sample_state("gravity", B, N, seed=10 + i)
    his, x, v, ea = synth.segno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), row, col)
    batches.append(dict(his=his, x=x, v=v, ea=ea, target=x + 0.05 * torch.randn_like(x)))
m = nb.SEGNO(in_node_nf=1, in_edge_nf=2, hidden_nf=64, device=dev, n_layers=8, recurrent=True)
opt = nb.FlatAdam(m.parameters(), lr=1e-3)


def loss_fn(his, x, v, ea, target):
    xo, _, _ = m(his, x, [row, col], v, ea, T=T)
    return nb.trajectory_mse(xo, target, 1)[0]


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


g = nb.GraphedStep(loss_fn, batches[0], opt)
step = lambda i: g(**batches[i % 4])
for i in range(5):
    step(i)
ms_train = timed(step, K)
with torch.no_grad():
    inf = lambda i: m(batches[i % 4]["his"], batches[i % 4]["x"], [row, col], batches[i % 4]["v"], batches[i % 4]["ea"], T=T)
    for i in range(3):
        inf(i)
    ms_inf = timed(inf, K)
lib = nb.load_library()


def eager(i):
    opt.zero_grad(set_to_none=True)
    loss_fn(**batches[i % 4]).backward()
    opt.step()


lib.nb_profile_enable(1)
timed(eager, 10)
sm_, sc_ = (ctypes.c_double * 8)(), (ctypes.c_longlong * 8)()
lib.nb_profile_read(sm_, sc_)
lib.nb_profile_enable(0)
cat = {"edge_bwd": 1, "segno_fused_fwd": 5, "node": 2, "wgrad64": 3}
print(json.dumps({"train_ms": round(ms_train, 4), "train_traj_per_s": round(B / ms_train * 1e3), "infer_ms": round(ms_inf, 4),
                  "infer_traj_per_s": round(B / ms_inf * 1e3), "launches_per_step": g.launches_per_replay,
                  "us_per_launch": {c: (round(sm_[i] / sc_[i] * 1e3, 1), int(sc_[i]) // 10) for c, i in cat.items() if sc_[i]}}))
