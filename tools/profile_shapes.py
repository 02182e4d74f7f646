#!/usr/bin/env python
"""A few eager training steps at the shapes bench.py's side legs use, for profiler runs (tools/profile_round.sh):
    segno    SEGNO gravity, N=20, 10 sub-steps, B=256   (BASELINE.json configs[3])
    egno100  EGNO charged, N=100, T=10, L=4, B=64        (BASELINE.json configs[4], per-GPU share at 8 GPUs)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import no_node_comparison_b200 as nb
from no_node_comparison_b200 import synth

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "segno"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if what == "segno":
    B, N, T = 256, 20, 10
    s = synth.sample_state("gravity", B, N, seed=1)
    row, col = synth.canonical_edges(B, N, dev)
    his, x, v, ea = synth.segno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), row, col)
    m = nb.SEGNO(in_node_nf=1, in_edge_nf=2, hidden_nf=64, device=dev, n_layers=8, recurrent=True)
    opt = nb.FlatAdam(m.parameters(), lr=1e-3)
    tgt = x + 0.05 * torch.randn_like(x)
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        xo, _, _ = m(his, x, [row, col], v, ea, T=T)
        nb.trajectory_mse(xo, tgt, 1)[0].backward()
        opt.step()
else:
    B, N, T, L = 64, 100, 10, 4
    s = synth.sample_state("charged", B, N, seed=1)
    row, col = synth.canonical_edges(B, N, dev)
    x, nodes, ea, v, lm = synth.egno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), row, col)
    m = nb.EGNO(n_layers=L, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T, device=dev)
    opt = nb.FlatAdam(m.parameters(), lr=1e-4)
    t_out = torch.arange(1, T + 1, device=dev)[None].repeat(B, 1)
    tgt = x.repeat(T, 1)
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        xo, _, _ = m(x, nodes, [row, col], ea, v=v, loc_mean=lm, timesteps_out=t_out)
        nb.trajectory_mse(xo, tgt, T)[0].backward()
        opt.step()
torch.cuda.synchronize()
print("ok", what)
