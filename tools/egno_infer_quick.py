#!/usr/bin/env python
"""EGNO inference timing at the configs[2] shape (N=20, T=10, L=4, B=256): no_grad forward, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import no_node_comparison_b200 as nb
from no_node_comparison_b200 import synth

dev = torch.device("cuda:0")
B, N, T, L = 256, 20, 10, 4
row, col = synth.canonical_edges(B, N, dev)
bs = []
for i in range(4):
    s = synth.sample_state("charged", B, N, seed=i)
    bs.append([t.to(dev) for t in synth.egno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), row, col)])
m = nb.EGNO(n_layers=L, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T, device=dev)
t_out = torch.arange(1, T + 1, device=dev)[None].repeat(B, 1)
with torch.no_grad():
    f = lambda i: m(bs[i % 4][0], bs[i % 4][1], [row, col], bs[i % 4][2], v=bs[i % 4][3], loc_mean=bs[i % 4][4], timesteps_out=t_out)
    for i in range(5):
        f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        f(i)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print({"infer_ms": round(ms, 4), "infer_traj_per_s": round(B / ms * 1e3)})
