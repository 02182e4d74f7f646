"""Diagnostic: how long must EGNO train on a small simulated set before a 20-call rollout stays bounded?"""
import sys, torch
sys.path.insert(0, '.')
import no_node_comparison_b200 as nb
from oracle import nbody_oracle as O
from tests.test_gpu_curves import _frames
DEV = torch.device('cuda:0')
N, T, L, B, CALLS, START = 20, 10, 4, 16, 20, 30
NTR = int(sys.argv[1]) if len(sys.argv) > 1 else 32
loc, vel, q = _frames("charged", NTR + 16, N, START + T * CALLS + 1, seed=43)
row, col = O.canonical_edges(B, N)
edges_d = [row.to(DEV), col.to(DEV)]
t_out = torch.arange(1, T + 1)[None].repeat(B, 1).to(DEV)
def batch(lo, f0=START):
    s = slice(lo, lo + B)
    x, v, ea, nodes, mean = O.egno_features(loc[s, f0], vel[s, f0], q[s], row, col)
    tgt = loc[s, f0 + 1:f0 + T + 1].permute(0, 2, 1, 3).reshape(B * N, T, 3)
    return {k: t.to(DEV) for k, t in dict(x=x, v=v, ea=ea, nodes=nodes, mean=mean, tgt=tgt).items()}
train = [batch(lo, f0) for lo in range(0, NTR, 16) for f0 in (START, START + 50, START + 100, START + 150)]
torch.manual_seed(1)
m = nb.EGNO(n_layers=L, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T, device=DEV)
opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=1e-12)
s = slice(NTR, NTR + B)
l0, v0, qq = loc[s, START].reshape(-1, 3).to(DEV), vel[s, START].reshape(-1, 3).to(DEV), q[s].reshape(-1, 1).to(DEV)
truth = loc[s, START + 1:START + T * CALLS + 1].transpose(0, 1)
for step in range(4001):
    b = train[step % len(train)]
    opt.zero_grad(set_to_none=True)
    xo, _, _ = m(b["x"], b["nodes"], edges_d, b["ea"], v=b["v"], loc_mean=b["mean"], timesteps_out=t_out)
    loss = nb.trajectory_mse(xo, b["tgt"], T)[0]
    loss.backward(); opt.step()
    if step in (50, 650, 1000, 2000, 4000):
        preds, e_last, e_all = nb.egno_rollout(m, l0, v0, qq, edges_d, N, traj_len=CALLS, dataset="charged")
        mx = preds.view(CALLS, T, -1).abs().amax(dim=(1, 2))
        mse = ((preds.cpu().view(CALLS * T, B, N, 3) - truth) ** 2).mean(dim=(1, 2, 3)).view(CALLS, T).mean(1)
        print(step, f"loss {float(loss):.4f}", "max|x| per call", [f"{v:.1f}" for v in mx.tolist()])
        print("    mse per call", [f"{v:.2f}" for v in mse.tolist()])
