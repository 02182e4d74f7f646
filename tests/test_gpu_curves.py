"""Curve-level parity (north_star: "test MSE and energy-conservation curves matching"; VERDICT r01 missing #2, next #3).

On simulated trajectories (the device simulators of this repo, pinned to the reference's synthetic_sim.py by
tests/test_sim.py) the CUDA modules and the *oracle model* on the CPU

  1. take the same 50 Adam steps (torch.optim.Adam on both, lr from model_confs.yaml) — loss curves and held-out test MSE
     must agree (test MSE within 1 %);
  2. roll out 20 autoregressive calls from the same trained weights (EGNO/main_simulation_simple_no.py:342-384,
     SEGNO/train_nbody.py:200-236) — per-call MSE against the simulated truth, the conserved-energy curve
     (utils.py:126-219) and the energy-drift statistics must agree, within a tolerance that grows with the horizon.
     A rollout is an iterated map: any fp32-level difference is amplified call after call (x2.2 per call for the
     20-body gravity model below; the EGNO callers feed back a velocity no loss term supervises, and a freshly trained
     model overflows after a few calls on every path).  The growth is therefore MEASURED, not guessed: the oracle
     rollout is repeated with its state perturbed by 1e-5 (relative; the size of one call's CUDA-vs-oracle difference)
     before every call, and the CUDA rollout must stay within 10x of that envelope per call (floor 1e-4 for positions,
     3e-4 for the MSE and energy curves), its statistics within 5x; the comparison covers the calls before the oracle
     model overflows, and the CUDA rollout must overflow at the same call.

Shapes are BASELINE.json configs[2] (EGNO, charged, N=20, T=10, L=4) and configs[3] (SEGNO, gravity, N=20, 10 sub-steps).
Also here: the full-batch (B=256) and 100-body oracle comparisons the property tests of test_gpu_parity.py do not make.
"""
from __future__ import annotations

import pytest
import torch

import no_node_comparison_b200 as nb
from oracle import nbody_oracle as O
from tests import dataset_factory as DF
from tests.helpers import param_grads_within_kink_budget as _param_grads_within_kink_budget, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _frames(kind, n_traj, N, n_frames, seed):
    """-> loc, vel [S, F, N, 3] fp32 (host), charges / masses [S, N, 1]"""
    loc, vel, _, q = DF.simulate_split(kind, N, n_traj, (n_frames + 1) * 100, 100, seed, device=DEV)
    if kind == "charged":                      # generate_dataset.py layout [S, F, 3, N]
        loc, vel = loc.transpose(0, 1, 3, 2), vel.transpose(0, 1, 3, 2)
    t = lambda a: torch.tensor(a, dtype=torch.float32)
    return t(loc), t(vel), t(q)


def _curve_report(name, a, b):
    r = [abs(x - y) / max(abs(y), 1e-12) for x, y in zip(a, b)]
    print(f"{name}: max rel diff {max(r):.2e} (first {r[0]:.2e}, last {r[-1]:.2e})")
    return r


# ------------------------------------------------------------------------------------------------ EGNO, configs[2]
def test_egno_training_and_rollout_curves_match_the_oracle_model():
    N, T, L, B, CALLS, START = 20, 10, 4, 16, 20, 30
    NTR = 128
    loc, vel, q = _frames("charged", NTR + B, N, START + T * CALLS + 1, seed=43)
    row, col = O.canonical_edges(B, N)
    edges_d = [row.to(DEV), col.to(DEV)]
    t_out = torch.arange(1, T + 1)[None].repeat(B, 1)

    def batch(lo, f0=START):
        s = slice(lo, lo + B)
        x, v, ea, nodes, mean = O.egno_features(loc[s, f0], vel[s, f0], q[s], row, col)
        tgt = loc[s, f0 + 1:f0 + T + 1].permute(0, 2, 1, 3).reshape(B * N, T, 3)      # [BN, T, 3]
        return dict(x=x, v=v, ea=ea, nodes=nodes, mean=mean, tgt=tgt)

    # windows starting at several frames of every training trajectory: the rollout must stay on the data distribution
    train = [batch(lo, f0) for lo in range(0, NTR, B) for f0 in (START, START + 50, START + 100, START + 150)]
    test = batch(NTR)
    torch.manual_seed(1)
    m = nb.EGNO(n_layers=L, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T, device=DEV)
    p = {k: t.detach().cpu().clone().requires_grad_(True) for k, t in m.named_parameters()}
    opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=1e-12)
    opt_r = torch.optim.Adam(list(p.values()), lr=5e-4, weight_decay=1e-12)

    def cuda_loss(b):
        xo, _, _ = m(b["x"].to(DEV), b["nodes"].to(DEV), edges_d, b["ea"].to(DEV), v=b["v"].to(DEV), loc_mean=b["mean"].to(DEV),
                     timesteps_out=t_out.to(DEV))
        return nb.trajectory_mse(xo, b["tgt"].to(DEV), T)[0]

    def oracle_loss(b):
        xo, _, _ = O.egno_forward(p, b["x"], b["nodes"], row, col, b["ea"], b["v"], b["mean"], t_out, n_layers=L, num_timesteps=T)
        return O.trajectory_mse(xo, b["tgt"])[0]

    ours, theirs = [], []
    for step in range(50):      # two batches of the reference's own window (frame 30 -> 31..40), alternating
        b = train[(step % 2) * 4]
        opt.zero_grad(set_to_none=True)
        lo = cuda_loss(b)
        lo.backward()
        opt.step()
        ours.append(float(lo.detach()))
        opt_r.zero_grad(set_to_none=True)
        lr_ = oracle_loss(b)
        lr_.backward()
        opt_r.step()
        theirs.append(float(lr_.detach()))
    r = _curve_report("EGNO train loss, 50 Adam steps", ours, theirs)
    assert max(r[:5]) < 1e-4 and max(r) < 5e-3, r
    with torch.no_grad():
        mse_o, mse_r = float(cuda_loss(test)), float(oracle_loss(test))
    print(f"EGNO held-out test MSE after training: cuda {mse_o:.6e} oracle {mse_r:.6e}")
    assert abs(mse_o - mse_r) < 1e-2 * mse_r

    # ---- 20-call rollout from the SAME weights on both paths (1950 more CUDA-only steps first: a model that has seen 50
    # steps leaves the data distribution within two calls and overflows, on both paths)
    train_d = [{k: t.to(DEV) for k, t in b.items()} for b in train]
    for step in range(50, 2000):
        opt.zero_grad(set_to_none=True)
        cuda_loss(train_d[step % len(train)]).backward()
        opt.step()
    w = {k: t.detach().cpu().clone() for k, t in m.named_parameters()}
    s = slice(NTR, NTR + B)
    truth = loc[s, START + 1:START + T * CALLS + 1]                       # [B, CALLS*T, N, 3]
    l0, v0, qq = loc[s, START].reshape(-1, 3), vel[s, START].reshape(-1, 3), q[s].reshape(-1, 1)
    preds, e_last, e_all = nb.egno_rollout(m, l0.to(DEV), v0.to(DEV), qq.to(DEV), edges_d, N, traj_len=CALLS, dataset="charged")
    preds, e_all = preds.cpu().view(CALLS * T, B, N, 3), e_all.cpu()

    @torch.no_grad()
    def oracle_rollout(lc, vc, eps=0.0, seed=0):
        ref_p, ref_e = [], []
        for k in range(CALLS):
            lc, vc = _perturbed(lc, 2 * k + 100 * seed, eps), _perturbed(vc, 2 * k + 1 + 100 * seed, eps)
            x, v, ea, nodes, mean = O.egno_features(lc, vc, q[s], row, col)
            xo, vo, _ = O.egno_forward(w, x, nodes, row, col, ea, v, mean, t_out, n_layers=L, num_timesteps=T)
            xo, vo = xo.view(T, B, N, 3), vo.view(T, B, N, 3)
            ref_p.append(xo)
            ref_e += [O.energy_charged(xo[t], vo[t], q[s]) for t in range(T)]
            lc, vc = xo[T - 1], vo[T - 1]
        return torch.cat(ref_p), torch.stack(ref_e)                      # [CALLS*T, B, N, 3], [CALLS*T, B]

    ref = oracle_rollout(l0.view(B, N, 3), v0.view(B, N, 3))
    pert = [oracle_rollout(l0.view(B, N, 3), v0.view(B, N, 3), eps=1e-5, seed=k) for k in range(3)]
    _check_rollout("EGNO", preds, e_all, ref, pert, truth.transpose(0, 1), CALLS, T, min_calls=2)


def _perturbed(t, seed, eps):
    return t if eps == 0.0 else t * (1.0 + eps * torch.randn(t.shape, generator=torch.Generator().manual_seed(seed)))


def _check_rollout(name, preds, en, ref, perts, truth, calls, per_call, min_calls):
    """preds / truth [F, B, N, 3], en [F, B]: the CUDA rollout; ref = (predictions, energies) of the oracle model, perts =
    the same from three runs whose state is perturbed by 1e-5 before every call (the measured error-growth envelope: the
    largest of the three, never shrinking with the horizon); F = calls * per_call frames."""
    ref_p, ref_en = ref
    # valid horizon: the leading calls over which the oracle model (all runs) stays within 1000x the largest true
    # coordinate; a learned simulator that leaves the data distribution overflows within two or three calls, on every path.
    # Curves are compared over the valid horizon; the CUDA rollout must leave it with the oracle.
    bound = 1e3 * truth.abs().amax().item()
    amax = lambda p: torch.nan_to_num(p.abs(), nan=float("inf")).amax(dim=(1, 2, 3)).view(calls, per_call).amax(1)
    ok = amax(ref_p) <= bound
    for pp, _ in perts:
        ok &= amax(pp) <= bound
    K = int(ok.long().cumprod(0).sum())
    print(f"{name} rollout: valid horizon {K} of {calls} calls (bound {bound:.1f}); max|x| per call, oracle {[f'{v:.3g}' for v in amax(ref_p).tolist()]}")
    assert K >= min_calls, K
    assert bool((amax(preds)[:K] <= bound).all()), amax(preds).tolist()
    if K < calls:
        assert bool((amax(preds)[K:min(K + 3, calls)] > bound).any()), amax(preds).tolist()
        calls = K
        cut = lambda t: t[:K * per_call]
        preds, en, truth, ref_p, ref_en = map(cut, (preds, en, truth, ref_p, ref_en))
        perts = [(cut(a), cut(b)) for a, b in perts]
    scale = ref_p.abs().amax().item()
    esc = ref_en.abs().mean()
    per = lambda a: (a - ref_p).abs().amax(dim=(1, 2, 3)).view(calls, per_call).amax(1) / scale
    mse = lambda p: ((p - truth) ** 2).mean(dim=(1, 2, 3)).view(calls, per_call).mean(1)
    mse_r = mse(ref_p)
    mcurve = lambda p: (mse(p) - mse_r).abs() / mse_r
    ecurve = lambda e: ((e.mean(1) - ref_en.mean(1)).abs() / esc).view(calls, per_call).amax(1)   # batch-mean energy per frame
    envelope = lambda f, k: torch.cummax(torch.stack([f(pt[k]) for pt in perts]).amax(0), 0).values
    d, rm, de = per(preds), mcurve(preds), ecurve(en)
    d_env, rm_env, de_env = envelope(per, 0), envelope(mcurve, 0), envelope(ecurve, 1)
    fmt = lambda t: [f"{v:.1e}" for v in t.tolist()]
    print(f"{name} rollout: test MSE vs truth per call (oracle)  {[f'{v:.2e}' for v in mse_r.tolist()]}")
    print(f"{name} rollout: prediction difference per call, cuda {fmt(d)}\n{'':>49}envelope {fmt(d_env)}")
    print(f"{name} rollout: test-MSE curve difference,      cuda {fmt(rm)}\n{'':>49}envelope {fmt(rm_env)}")
    print(f"{name} rollout: energy curve difference,        cuda {fmt(de)}\n{'':>49}envelope {fmt(de_env)}")
    for k in range(calls):
        assert d[k] < max(1e-4, 10 * d_env[k]), (name, "predictions", k, float(d[k]), float(d_env[k]))
        # scalar curves fluctuate more between chaotic realisations than the max norm: bound them by their own envelope
        # or by the prediction tolerance, whichever is larger
        assert rm[k] < max(3e-4, 10 * rm_env[k], 10 * d_env[k]), (name, "test MSE", k, float(rm[k]), float(rm_env[k]))
        assert de[k] < max(3e-4, 10 * de_env[k], 10 * d_env[k]), (name, "energy", k, float(de[k]), float(de_env[k]))
    assert d[0] < 1e-4 and rm[0] < 3e-4 and de[0] < 3e-4                # first call: plain fp32 tolerance
    # statistics over the whole valid horizon: mean test MSE and mean energy drift (utils.compute_energy_drift)
    drift = lambda e: ((e - e[0:1]).abs() / e[0:1].abs().clamp_min(1e-3 * esc)).mean()
    st = lambda p, e: torch.stack([mse(p).mean(), drift(e)])
    s_o, s_r = st(preds, en), st(ref_p, ref_en)
    s_env = torch.stack([(st(a, b) - s_r).abs() for a, b in perts]).amax(0)
    print(f"{name} rollout: [mean test MSE, mean energy drift] cuda {s_o.tolist()} oracle {s_r.tolist()} envelope {s_env.tolist()}")
    tol = torch.maximum(5 * s_env, 1e-2 * s_r.abs())
    assert ((s_o - s_r).abs() <= tol).all(), (s_o.tolist(), s_r.tolist(), tol.tolist())


# ------------------------------------------------------------------------------------------------ SEGNO, configs[3]
def test_segno_training_and_rollout_curves_match_the_oracle_model():
    N, T, B, CALLS, STRIDE = 20, 10, 16, 20, 2
    NTR = 128
    loc, vel, mass = _frames("gravity", NTR + B, N, STRIDE * CALLS + 1, seed=47)
    row, col = O.canonical_edges(B, N)
    edges_d = [row.to(DEV), col.to(DEV)]

    def batch(lo, f0=0):
        s = slice(lo, lo + B)
        his, x, v, ea = O.segno_features(loc[s, f0], vel[s, f0], mass[s], row, col)
        return dict(his=his, x=x, v=v, ea=ea, tgt=loc[s, f0 + STRIDE].reshape(-1, 3))

    train = [batch(lo, f0) for lo in range(0, NTR, B) for f0 in (0, 10, 20, 30)]
    test = batch(NTR)
    torch.manual_seed(1)
    m = nb.SEGNO(in_node_nf=1, in_edge_nf=2, hidden_nf=64, device=DEV, n_layers=8, recurrent=True)
    p = {k: t.detach().cpu().clone().requires_grad_(True) for k, t in m.named_parameters()}
    live = [k for k in p if "coord_mlp_vel" not in k]
    opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=1e-12)
    opt_r = torch.optim.Adam([p[k] for k in live], lr=5e-4, weight_decay=1e-12)
    mse = torch.nn.MSELoss()

    def cuda_loss(b):
        xo, _, _ = m(b["his"].to(DEV), b["x"].to(DEV), edges_d, b["v"].to(DEV), b["ea"].to(DEV), T=T)
        return mse(xo, b["tgt"].to(DEV))

    def oracle_loss(b):
        xo, _, _ = O.segno_forward(p, b["his"], b["x"], row, col, b["v"], b["ea"], T)
        return mse(xo, b["tgt"])

    ours, theirs = [], []
    for step in range(50):
        b = train[step % len(train)]
        opt.zero_grad(set_to_none=True)
        lo = cuda_loss(b)
        lo.backward()
        opt.step()
        ours.append(float(lo.detach()))
        opt_r.zero_grad(set_to_none=True)
        lr_ = oracle_loss(b)
        lr_.backward()
        opt_r.step()
        theirs.append(float(lr_.detach()))
    r = _curve_report("SEGNO train loss, 50 Adam steps", ours, theirs)
    assert max(r[:5]) < 1e-4 and max(r) < 5e-3, r
    with torch.no_grad():
        mse_o, mse_r = float(cuda_loss(test)), float(oracle_loss(test))
    print(f"SEGNO held-out test MSE after training: cuda {mse_o:.6e} oracle {mse_r:.6e}")
    assert abs(mse_o - mse_r) < 1e-2 * mse_r

    train_d = [{k: t.to(DEV) for k, t in b.items()} for b in train]
    for step in range(50, 2000):
        opt.zero_grad(set_to_none=True)
        cuda_loss(train_d[step % len(train)]).backward()
        opt.step()
    w = {k: t.detach().cpu().clone() for k, t in m.named_parameters()}
    s = slice(NTR, NTR + B)
    truth = loc[s, STRIDE::STRIDE][:, :CALLS].transpose(0, 1)            # [CALLS, B, N, 3]
    l0, v0, mm = loc[s, 0].reshape(-1, 3), vel[s, 0].reshape(-1, 3), mass[s].reshape(-1, 1)
    preds, en = nb.segno_rollout(m, l0.to(DEV), v0.to(DEV), mm.to(DEV), edges_d, N, traj_len=CALLS, num_steps=T, dataset="gravity")

    @torch.no_grad()
    def oracle_rollout(lc, vc, eps=0.0, seed=0):
        ref_p, ref_e = [], []
        for k in range(CALLS):
            lc, vc = _perturbed(lc, 2 * k + 100 * seed, eps), _perturbed(vc, 2 * k + 1 + 100 * seed, eps)
            his, x, v, ea = O.segno_features(lc, vc, mass[s], row, col)
            xo, _, vo = O.segno_forward(w, his, x, row, col, v, ea, T)
            lc, vc = xo.view(B, N, 3), vo.view(B, N, 3)
            ref_p.append(lc)
            ref_e.append(O.energy_gravity(lc, vc, mass[s]))
        return torch.stack(ref_p), torch.stack(ref_e)

    ref = oracle_rollout(l0.view(B, N, 3), v0.view(B, N, 3))
    pert = [oracle_rollout(l0.view(B, N, 3), v0.view(B, N, 3), eps=1e-5, seed=k) for k in range(3)]
    _check_rollout("SEGNO", preds.cpu().view(CALLS, B, N, 3), en.cpu(), ref, pert, truth, CALLS, 1, min_calls=20)


# ------------------------------------------------------------------------------------------------ full-size oracle comparisons
def _egno_vs_oracle(B, N, T, L, seed, cond_factor=0.0):
    from no_node_comparison_b200 import synth

    s = synth.sample_state("charged", B, N, seed)
    row, col = synth.canonical_edges(B, N)
    x, nodes, ea, v, lm = synth.egno_features(s["loc"], s["vel"], s["charges"], row, col)
    t_out = torch.arange(1, T + 1)[None].repeat(B, 1)
    torch.manual_seed(1)
    m = nb.EGNO(n_layers=L, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T, device=DEV)
    p = {k: t.detach().cpu().clone().requires_grad_(True) for k, t in m.named_parameters()}
    xg = x.to(DEV).requires_grad_(True)
    xo, vo, ho = m(xg, nodes.to(DEV), [row.to(DEV), col.to(DEV)], ea.to(DEV), v=v.to(DEV), loc_mean=lm.to(DEV),
                   timesteps_out=t_out.to(DEV))
    gen = torch.Generator().manual_seed(3)
    Gx, Gh = torch.randn(xo.shape, generator=gen), torch.randn(ho.shape, generator=gen) * 0.05
    ((xo * Gx.to(DEV)).sum() + (ho * Gh.to(DEV)).sum()).backward()
    xr = x.clone().requires_grad_(True)
    xo_r, vo_r, ho_r = O.egno_forward(p, xr, nodes, row, col, ea, v, lm, t_out, n_layers=L, num_timesteps=T)
    ((xo_r * Gx).sum() + (ho_r * Gh).sum()).backward()
    errs = dict(x=rel_err(xo.cpu(), xo_r.detach()), v=rel_err(vo.cpu(), vo_r.detach()), h=rel_err(ho.cpu(), ho_r.detach()),
                gx=rel_err(xg.grad.cpu(), xr.grad))
    print(f"EGNO B={B} N={N} T={T} L={L} vs oracle:", {k: f"{e:.1e}" for k, e in errs.items()})
    # conditioning of the case: the fp32 oracle against the same oracle in float64
    with torch.no_grad():
        pd = {k: t.detach().double() for k, t in p.items()}
        xo_d, vo_d, ho_d = O.egno_forward(pd, x.double(), nodes.double(), row, col, ea.double(), v.double(), lm.double(), t_out,
                                          n_layers=L, num_timesteps=T)
    noise = max(rel_err(xo_r.detach(), xo_d), rel_err(vo_r.detach(), vo_d), rel_err(ho_r.detach(), ho_d))
    print(f"   fp32 oracle vs float64 oracle: {noise:.1e}; displacement max|x_out - x_in| = "
          f"{(xo_d - x.double().repeat(T, 1)).abs().max():.2f} at max|x_in| = {x.abs().max():.2f}")
    tol_out = max(1e-4, cond_factor * noise)
    assert max(errs["x"], errs["v"], errs["h"]) < tol_out and errs["gx"] < 1e-3, (errs, tol_out)
    return m, p


def test_egno_full_batch_config3_against_the_oracle():
    """BASELINE.json configs[2] at the benchmark batch: B=256, N=20, T=10, L=4 — outputs, input and parameter gradients."""
    m, p = _egno_vs_oracle(256, 20, 10, 4, seed=5)
    _param_grads_within_kink_budget(m, p, n_rows=10 * 256 * 20)


def test_egno_100_body_config5_against_the_oracle():
    """BASELINE.json configs[4] shape (blocked receiver x sender walk): N=100, B=4, T=10, L=4.

    With random-initial weights this case is badly conditioned: messages are SUMMED over 99 neighbours, four layers move
    the particles by 35 length units from inputs of size 3, and PyTorch's own fp32 evaluation sits 1e-5 from float64 (2e-7
    at N=20).  The 1e-4 bound of BASELINE.json therefore holds per layer (L=1 below: 2e-6 / 1e-5), and the four-layer
    outputs are held to 30x the reference's own fp32-vs-float64 distance (observed 16x: 1.7e-4)."""
    _egno_vs_oracle(4, 100, 10, 1, seed=6)
    m, p = _egno_vs_oracle(4, 100, 10, 4, seed=6, cond_factor=30.0)
    _param_grads_within_kink_budget(m, p, n_rows=10 * 4 * 100)
