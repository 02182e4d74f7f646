#!/usr/bin/env python
"""Data-parallel check on N GPUs (torchrun): the peer-memory path (all-reduce fused into Adam, nb_adam_step_peers) against
the NCCL flat-bucket all-reduce: parameters after a few steps (eager and graph replay) and the step time of each.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import no_node_comparison_b200 as nb
from no_node_comparison_b200 import synth
from no_node_comparison_b200.dataparallel import init_from_env, broadcast_parameters

rank, local, world = init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
B, N, T, L = 256, 20, 10, 4
row, col = synth.canonical_edges(B, N, dev)
s = synth.sample_state("charged", B, N, seed=100 + rank)
x, nodes, ea, v, lm = synth.egno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), row, col)
tgt = x.repeat(T, 1) + 0.05 * torch.randn(T * B * N, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
t_out = torch.arange(1, T + 1, device=dev)[None].repeat(B, 1)
ins = dict(x=x, nodes=nodes, ea=ea, v=v, lm=lm, tgt=tgt)
res = {}
for mode in ("nccl", "peer"):
    torch.manual_seed(1)
    m = nb.EGNO(n_layers=L, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T, device=dev)
    broadcast_parameters(m)
    m.enable_data_parallel(peer_memory=(mode == "peer"))
    opt = nb.FlatAdam(m.parameters(), lr=1e-3, peer_bucket=m.peer_bucket)

    def fn(x, nodes, ea, v, lm, tgt):
        xo, _, _ = m(x, nodes, [row, col], ea, v=v, loc_mean=lm, timesteps_out=t_out)
        return nb.trajectory_mse(xo, tgt, T)[0]

    for _ in range(3):      # eager
        opt.zero_grad(set_to_none=True)
        fn(**ins).backward()
        opt.step()
    p_eager = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone()
    g = nb.GraphedStep(fn, ins, opt, warmup=2)
    for _ in range(5):
        g(**ins)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        g.graph.replay()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 100], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    pf = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone()
    # replicas identical?
    ref = pf.clone(); dist.broadcast(ref, 0)
    same = torch.tensor([float(torch.equal(ref, pf))], device=dev); dist.all_reduce(same, op=dist.ReduceOp.MIN)
    res[mode] = (p_eager, pf, float(ms), bool(same.item()))
    if rank == 0:
        print(f"{mode}: {float(ms):.4f} ms/step (max over {world} ranks), replicas identical: {bool(same.item())}", flush=True)
d_e = (res["nccl"][0] - res["peer"][0]).abs().max().item()
d_f = (res["nccl"][1] - res["peer"][1]).abs().max().item() / res["nccl"][1].abs().max().item()
if rank == 0:
    print(f"max |param(nccl) - param(peer)| after 3 eager steps: {d_e:.3e}; relative after the graph steps: {d_f:.3e}")
    print(f"step time: nccl {res['nccl'][2]:.4f} ms, peer {res['peer'][2]:.4f} ms")
torch.cuda.synchronize(); dist.barrier(); os._exit(0)
