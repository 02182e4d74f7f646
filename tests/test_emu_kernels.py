"""Kernel logic on the CPU: the *same* csrc/*.cu sources compiled for the host CUDA emulator
(tests/emu) must reproduce the reference's golden outputs and gradients, and the oracle on edge cases.
This is test infrastructure — the product library is the nvcc build and has no CPU path."""
import numpy as np
import pytest
import torch

from oracle import nbody_oracle as O
from tests import emu_harness as E
from tests.helpers import (load_case, egno_inputs_from_case, egno_multi_inputs_from_case, segno_inputs_from_case, rel_err,
                           EGNO_MULTI_CASES)

TOL = 2e-5  # fp32 kernels vs fp32 reference: summation-order noise only


def _check_grads(r, w, g, order):
    off = 0
    for k in order:
        n = w[k].numel()
        got = torch.tensor(r["grad_params"][off:off + n]).view_as(w[k])
        assert rel_err(got, g[k]) < TOL, k
        off += n
    assert off == r["grad_params"].size


@pytest.mark.parametrize("name", ["egno_n5_t8", "egno_n5_t6_m4", "egno_n7_t10_m5"])
def test_egno_emulated_kernels_match_reference_golden(name):
    d, w, g = load_case(name)
    c = egno_inputs_from_case(d)
    order = list(w.keys())
    params = E.flat_params({k: v.numpy() for k, v in w.items()}, order)
    cfg = dict(B=c["B"], N=c["n"], T=c["T"], n_layers=c["L"], num_modes=c["modes"], in_node_nf=2, in_edge_nf=2,
               time_emb_dim=32, use_time_conv=1)
    r = E.egno_run(cfg, params, c["x"].numpy(), c["nodes"].numpy(), c["edge_attr"].numpy(), c["v"].numpy(),
                   c["loc_mean"].numpy(), c["t_out"].numpy(), d["Gx"], d["Gv"], d["Gh"])
    for k in ("x_out", "v_out", "h_out", "gx_in", "gv_in"):
        assert rel_err(torch.tensor(r[k]), torch.tensor(d[k])) < TOL, k
    _check_grads(r, w, g, order)


@pytest.mark.parametrize("name", EGNO_MULTI_CASES)
def test_egno_multi_input_emulated_kernels_match_reference_golden(name):
    """num_inputs > 1 (egno.py:42-47,59-61,68-70,79-86) and per-trajectory output times (varDT) against the reference."""
    d, w, g = load_case(name)
    c = egno_multi_inputs_from_case(d)
    order = list(w.keys())
    params = E.flat_params({k: v.numpy() for k, v in w.items()}, order)
    cfg = dict(B=c["B"], N=c["n"], T=c["T"], n_layers=c["L"], num_modes=c["modes"], in_node_nf=2, in_edge_nf=2,
               time_emb_dim=32, use_time_conv=1, num_inputs=c["num_inputs"])
    r = E.egno_run(cfg, params, c["x"].numpy(), c["nodes"].numpy(), c["edge_attr"].numpy(), c["v"].numpy(),
                   c["loc_mean"].numpy(), c["t_out"].numpy(), d["Gx"], d["Gv"], d["Gh"], t_in=c["t_in"].numpy())
    for k in ("x_out", "v_out", "h_out", "gx_in", "gv_in"):
        assert rel_err(torch.tensor(r[k]), torch.tensor(d[k])) < TOL, k
    _check_grads(r, w, g, order)


@pytest.mark.parametrize("name", ["segno_n5_t10", "segno_n20_t10_gravity"])
def test_segno_emulated_kernels_match_reference_golden(name):
    d, w, g = load_case(name)
    c = segno_inputs_from_case(d)
    order = list(w.keys())
    params = E.flat_params({k: v.numpy() for k, v in w.items()}, order)
    cfg = dict(B=c["B"], N=c["n"], T=c["T"], in_node_nf=1, in_edge_nf=2, recurrent=1, coords_weight=1.0)
    r = E.segno_run(cfg, params, c["his"].numpy(), c["x"].numpy(), c["v"].numpy(), c["edge_attr"].numpy(), d["Gx"],
                    d["Gh"], d["Gv"])
    for k in ("x_out", "v_out", "h_out", "gx_in", "gv_in"):
        assert rel_err(torch.tensor(r[k]), torch.tensor(d[k])) < TOL, k
    _check_grads(r, w, g, order)


@pytest.mark.parametrize("B,N,T", [(1, 2, 3), (7, 3, 2), (3, 12, 2)])
def test_segno_emulated_edge_shapes_vs_oracle(B, N, T):
    """Ragged packing (units of several graphs, partial last unit), N=2, and tiles straddling receivers."""
    _, w, _ = load_case("segno_n5_t10")
    order = list(w.keys())
    gen = torch.Generator().manual_seed(B * 100 + N)
    # amplify phi_x's last layer (xavier gain 1e-3 in the reference) so the coordinate path is exercised
    w = {k: v.clone() for k, v in w.items()}
    w["module.coord_mlp.2.weight"] *= 300.0
    loc = torch.randn(B, N, 3, generator=gen)
    vel = torch.randn(B, N, 3, generator=gen) * 0.5
    q = torch.randint(0, 2, (B, N, 1), generator=gen).float() * 2 - 1
    row, col = O.canonical_edges(B, N)
    his, x, v, ea = O.segno_features(loc, vel, q, row, col)
    Gx, Gv, Gh = torch.randn(B * N, 3, generator=gen), torch.randn(B * N, 3, generator=gen), torch.randn(B * N, 64, generator=gen) * 0.1
    p = {k: t.clone().requires_grad_(True) for k, t in w.items()}
    xr, vr = x.clone().requires_grad_(True), v.clone().requires_grad_(True)
    xo, ho, vo = O.segno_forward(p, his, xr, row, col, vr, ea, T)
    ((xo * Gx).sum() + (vo * Gv).sum() + (ho * Gh).sum()).backward()
    params = E.flat_params({k: t.numpy() for k, t in w.items()}, order)
    cfg = dict(B=B, N=N, T=T, in_node_nf=1, in_edge_nf=2, recurrent=1, coords_weight=1.0)
    r = E.segno_run(cfg, params, his.numpy(), x.numpy(), v.numpy(), ea.numpy(), Gx.numpy(), Gh.numpy(), Gv.numpy())
    assert rel_err(torch.tensor(r["x_out"]), xo.detach()) < TOL
    assert rel_err(torch.tensor(r["h_out"]), ho.detach()) < TOL
    assert rel_err(torch.tensor(r["v_out"]), vo.detach()) < TOL
    assert rel_err(torch.tensor(r["gx_in"]), xr.grad) < TOL
    assert rel_err(torch.tensor(r["gv_in"]), vr.grad) < TOL
    g = {k: (p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])) for k in order}
    _check_grads(r, w, g, order)


def test_segno_per_edge_clamp_is_reproduced():
    """rij * c beyond +-100 must be clamped per edge BEFORE the mean (gcl.py:100), forward and backward."""
    _, w, _ = load_case("segno_n5_t10")
    order = list(w.keys())
    w = {k: v.clone() for k, v in w.items()}
    w["module.coord_mlp.2.weight"] *= 3.0e5          # push |rij * c| past 100 for many edges
    B, N, T = 2, 5, 2
    gen = torch.Generator().manual_seed(5)
    loc = torch.randn(B, N, 3, generator=gen) * 2
    vel = torch.randn(B, N, 3, generator=gen) * 0.5
    q = torch.randint(0, 2, (B, N, 1), generator=gen).float() * 2 - 1
    row, col = O.canonical_edges(B, N)
    his, x, v, ea = O.segno_features(loc, vel, q, row, col)
    p = {k: t.clone().requires_grad_(True) for k, t in w.items()}
    xr = x.clone().requires_grad_(True)
    xo, ho, vo = O.segno_forward(p, his, xr, row, col, v, ea, T)
    Gx = torch.randn(B * N, 3, generator=gen)
    (xo * Gx).sum().backward()
    params = E.flat_params({k: t.numpy() for k, t in w.items()}, order)
    cfg = dict(B=B, N=N, T=T, in_node_nf=1, in_edge_nf=2, recurrent=1, coords_weight=1.0)
    z3, z64 = np.zeros((B * N, 3), np.float32), np.zeros((B * N, 64), np.float32)
    r = E.segno_run(cfg, params, his.numpy(), x.numpy(), v.numpy(), ea.numpy(), Gx.numpy(), z64, z3)
    assert float(xo.detach().abs().max()) > 5.0       # the clamp really was active
    assert rel_err(torch.tensor(r["x_out"]), xo.detach()) < TOL
    assert rel_err(torch.tensor(r["gx_in"]), xr.grad) < 1e-4


def test_edge_index_check_kernel():
    import ctypes
    L = E.lib()
    B, N = 3, 4
    row, col = O.canonical_edges(B, N)
    row, col = row.numpy().copy(), col.numpy().copy()
    flag = np.zeros(1, np.int32)
    E.check(L.nb_check_canonical_edges(E.ptr(row), E.ptr(col), row.size, B, N, E.ptr(flag), None))
    assert flag[0] == 0
    col[7], col[8] = col[8], col[7]
    E.check(L.nb_check_canonical_edges(E.ptr(row), E.ptr(col), row.size, B, N, E.ptr(flag), None))
    assert flag[0] in (8, 9)
    assert L.nb_check_canonical_edges(E.ptr(row), E.ptr(col), row.size - 1, B, N, E.ptr(flag), None) < 0



@pytest.mark.parametrize("B,N", [(3, 5), (2, 20), (1, 2)])
def test_featurisation_and_energy_kernels_vs_oracle(B, N):
    """nb_nbody_features / nb_nbody_energy (the device-side prepare_inputs and conserved-energy evaluation) against
    the oracle's restatements of main_simulation_simple_no.py:326-338 and utils.py:126-144, :175-195."""
    import ctypes
    L = E.lib()
    g = torch.Generator().manual_seed(B * 100 + N)
    loc, vel = torch.randn(B, N, 3, generator=g), torch.randn(B, N, 3, generator=g)
    q = torch.randint(0, 2, (B, N, 1), generator=g).float() * 2 - 1
    row, col = O.canonical_edges(B, N)
    x, v, ea_r, nodes_r, lm_r = O.egno_features(loc, vel, q, row, col)
    bn, n_edges = B * N, B * N * (N - 1)
    nodes, mean, ea = np.zeros((bn, 2), np.float32), np.zeros((bn, 3), np.float32), np.zeros((n_edges, 2), np.float32)
    E.check(L.nb_nbody_features(B, N, 1, E.ptr(E.f32(x)), E.ptr(E.f32(v)), E.ptr(E.f32(q.reshape(-1))), None, E.ptr(nodes),
                                E.ptr(mean), E.ptr(ea), None))
    np.testing.assert_allclose(nodes, nodes_r.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(mean, lm_r.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ea, ea_r.numpy(), rtol=1e-6, atol=1e-6)
    his = np.zeros((bn, 1), np.float32)
    E.check(L.nb_nbody_features(B, N, 0, E.ptr(E.f32(x)), E.ptr(E.f32(v)), E.ptr(E.f32(q.reshape(-1))), None, E.ptr(his),
                                None, E.ptr(ea), None))
    np.testing.assert_allclose(his[:, 0], nodes_r[:, 0].numpy(), rtol=1e-6)
    # energies over F = 2 frames
    F = 2
    locs = torch.stack([loc, loc + 0.1 * torch.randn(B, N, 3, generator=g)]).reshape(F, bn, 3)
    vels = torch.stack([vel, 0.5 * vel]).reshape(F, bn, 3)
    for kind, fn, ch in ((0, O.energy_charged, q), (1, O.energy_gravity, 1.0 + 0.1 * torch.rand(B, N, 1, generator=g))):
        out = np.zeros((F, B), np.float32)
        E.check(L.nb_nbody_energy(kind, F, B, N, ctypes.c_float(1.0), E.ptr(E.f32(locs)), E.ptr(E.f32(vels)),
                                  E.ptr(E.f32(ch.reshape(-1))), E.ptr(out), None))
        ref = torch.stack([fn(locs[f].reshape(B, N, 3), vels[f].reshape(B, N, 3), ch) for f in range(F)]).numpy()
        np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-5)


def test_fused_adam_kernel_vs_torch_adam():
    """nb_adam_step (one launch over a flat buffer, device-side step counter) against torch.optim.Adam on the CPU."""
    import ctypes
    L = E.lib()
    g = torch.Generator().manual_seed(0)
    n = 1000
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2)
    p = p0.numpy().copy()
    m, v, step = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(1, np.float32)
    for it in range(5):
        grad = torch.randn(n, generator=g)
        ref.grad = grad.clone()
        opt.step()
        # the kernel multiplies the gradient by grad_scale first (data parallel: 1 / world over the summed bucket)
        E.check(L.nb_adam_step(n, E.ptr(p), E.ptr(E.f32(grad * 4.0)), E.ptr(m), E.ptr(v), E.ptr(step), 1, ctypes.c_double(3e-3),
                               ctypes.c_double(0.9), ctypes.c_double(0.99), ctypes.c_double(1e-8), ctypes.c_double(1e-2),
                               ctypes.c_double(0.25), None))
        np.testing.assert_allclose(p, ref.detach().numpy(), rtol=5e-7, atol=5e-8)
    assert step[0] == 5.0


def test_fused_adam_kernel_summing_peer_gradient_buffers():
    """nb_adam_step_peers (data parallel: the all-reduce is fused into the update kernel, which reads every rank's flat
    gradient buffer) against torch.optim.Adam fed with the mean of the per-rank gradients; three emulated ranks."""
    import ctypes
    L = E.lib()
    g = torch.Generator().manual_seed(1)
    n, world = 777, 3
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3)
    p = p0.numpy().copy()
    m, v, step = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(1, np.float32)
    for it in range(4):
        grads = [E.f32(torch.randn(n, generator=g)) for _ in range(world)]
        ref.grad = torch.tensor((grads[0] + grads[1]) + grads[2]) * np.float32(1.0 / world)     # rank order, like the kernel
        opt.step()
        peers = (ctypes.c_void_p * world)(*[E.ptr(a) for a in grads])
        E.check(L.nb_adam_step_peers(n, E.ptr(p), peers, world, E.ptr(m), E.ptr(v), E.ptr(step), 1, ctypes.c_double(2e-3),
                                     ctypes.c_double(0.9), ctypes.c_double(0.999), ctypes.c_double(1e-8), ctypes.c_double(1e-3),
                                     ctypes.c_double(1.0 / world), None))
        np.testing.assert_allclose(p, ref.detach().numpy(), rtol=5e-7, atol=5e-8)
    assert step[0] == 4.0
    assert L.nb_adam_step_peers(n, E.ptr(p), peers, 9, E.ptr(m), E.ptr(v), E.ptr(step), 1, ctypes.c_double(2e-3),
                                ctypes.c_double(0.9), ctypes.c_double(0.999), ctypes.c_double(1e-8), ctypes.c_double(1e-3),
                                ctypes.c_double(1.0), None) != 0        # more than 8 peers: refused


@pytest.mark.parametrize("T,R,layout,only_first", [(8, 25, 1, False), (10, 400, 0, False), (10, 40, 1, True), (1, 100, 0, False)])
def test_trajectory_mse_kernel_vs_oracle(T, R, layout, only_first):
    """nb_traj_mse (the callers' loss and its gradient) against the oracle's restatement of
    main_simulation_simple_no.py:268-276 and torch autograd."""
    L = E.lib()
    g = torch.Generator().manual_seed(T * 1000 + R)
    pred = torch.randn(T * R, 3, generator=g, requires_grad=True)
    tgt_bnt3 = torch.randn(R, T, 3, generator=g)
    loss_r, losses_r = O.trajectory_mse(pred, tgt_bnt3, only_first)
    loss_r.backward()
    tgt = tgt_bnt3 if layout == 1 else tgt_bnt3.transpose(0, 1).reshape(T * R, 3)
    losses, loss = np.zeros(T, np.float32), np.zeros(1, np.float32)
    grad = np.zeros((T * R, 3), np.float32)
    ws = np.zeros(int(L.nb_traj_mse_workspace_floats(T)), np.float32)
    E.check(L.nb_traj_mse(T, R, layout, int(only_first), E.ptr(E.f32(pred.detach())), E.ptr(E.f32(tgt)), E.ptr(losses),
                          E.ptr(loss), E.ptr(grad), E.ptr(ws), None))
    np.testing.assert_allclose(losses, losses_r.detach().numpy(), rtol=2e-6)
    np.testing.assert_allclose(loss[0], loss_r.item(), rtol=2e-6)
    np.testing.assert_allclose(grad, pred.grad.numpy(), rtol=1e-6, atol=1e-9)
    assert L.nb_traj_mse(0, R, layout, 0, E.ptr(grad), E.ptr(grad), E.ptr(losses), E.ptr(loss), None, E.ptr(ws), None) < 0


def _attn_params(g):
    W0 = torch.randn(64, 65, generator=g) * 0.2
    b0 = torch.randn(64, generator=g) * 0.1
    w2 = torch.randn(64, generator=g) * 0.3
    b2 = torch.randn(1, generator=g) * 0.1
    return W0, b0, w2, b2


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_segno_merge_kernels_vs_torch(mode):
    """nb_segno_merge_forward / _backward (multi-input SEGNO between its integration segments, model.py:82-90, 105-139:
    copy of an observed frame, 'sum', InvariantTemporalAttention + prepare_node_inputs) against torch autograd of the
    reference's formulas; the weight gradients are accumulated over two calls."""
    L = E.lib()
    g = torch.Generator().manual_seed(40 + mode)
    n, nf, frame = 37, 3, 1
    h_all = torch.randn(n, nf, 64, generator=g, requires_grad=True)
    x_all = torch.randn(n, nf, 3, generator=g, requires_grad=True)
    v_all = (torch.randn(n, nf, 3, generator=g) * 0.5).requires_grad_(True)
    h_int = torch.randn(n, 64, generator=g, requires_grad=True)
    x_int = torch.randn(n, 3, generator=g, requires_grad=True)
    v_int = (torch.randn(n, 3, generator=g) * 0.5).requires_grad_(True)
    W0, b0, w2, b2 = [t.requires_grad_(True) for t in _attn_params(g)]
    Gh, Gx, Gv = torch.randn(n, 64, generator=g), torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    ho, xo, vo = h_all[:, frame], x_all[:, frame], v_all[:, frame]
    if mode == 0:
        h_r, x_r, v_r = ho, xo, vo
    elif mode == 1:
        h_r, x_r, v_r = ho + h_int, xo + x_int, vo + v_int
    else:
        hs, xs, vs = torch.stack([ho, h_int], 1), torch.stack([xo, x_int], 1), torch.stack([vo, v_int], 1)
        feats = torch.cat([vs.norm(dim=-1, keepdim=True), hs], dim=-1)
        attn = (torch.tanh(feats @ W0.T + b0) @ w2[:, None] + b2).softmax(dim=1)
        x_r, v_r, h_r = (attn * xs).sum(1), (attn * vs).sum(1), (attn * hs).sum(1)
    ((h_r * Gh).sum() + (x_r * Gx).sum() + (v_r * Gv).sum()).backward()
    ap = E.f32(torch.cat([W0.detach().reshape(-1), b0.detach(), w2.detach(), b2.detach()]))
    arr = lambda t: E.f32(t.detach())
    h_out, x_out, v_out = np.zeros((n, 64), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    alpha = np.zeros((n, 2), np.float32)
    intp = (None, None, None) if mode == 0 else (E.ptr(arr(h_int)), E.ptr(arr(x_int)), E.ptr(arr(v_int)))
    E.check(L.nb_segno_merge_forward(mode, n, nf, frame, E.ptr(arr(h_all)), E.ptr(arr(x_all)), E.ptr(arr(v_all)), *intp,
                                     E.ptr(ap) if mode == 2 else None, E.ptr(h_out), E.ptr(x_out), E.ptr(v_out),
                                     E.ptr(alpha) if mode == 2 else None, None))
    np.testing.assert_allclose(h_out, h_r.detach().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(x_out, x_r.detach().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(v_out, v_r.detach().numpy(), rtol=2e-5, atol=2e-6)
    gha, gxa, gva = np.full((n, nf, 64), 7.0, np.float32), np.full((n, nf, 3), 7.0, np.float32), np.full((n, nf, 3), 7.0, np.float32)
    ghi, gxi, gvi = np.zeros((n, 64), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    g_attn = np.zeros(64 * 65 + 129, np.float32)
    ws = np.zeros(int(L.nb_segno_merge_backward_workspace_floats(n)), np.float32)
    gint = (None, None, None) if mode == 0 else (E.ptr(ghi), E.ptr(gxi), E.ptr(gvi))
    for acc in (0, 1):
        E.check(L.nb_segno_merge_backward(mode, n, nf, frame, E.ptr(arr(h_all)), E.ptr(arr(x_all)), E.ptr(arr(v_all)), *intp,
                                          E.ptr(ap) if mode == 2 else None, E.ptr(alpha) if mode == 2 else None, E.ptr(E.f32(Gh)),
                                          E.ptr(E.f32(Gx)), E.ptr(E.f32(Gv)), E.ptr(gha), E.ptr(gxa), E.ptr(gva), *gint,
                                          E.ptr(g_attn) if mode == 2 else None, acc, E.ptr(ws) if mode == 2 else None, None))
    tol = dict(rtol=5e-5, atol=5e-6)
    np.testing.assert_allclose(gha[:, frame], h_all.grad[:, frame].numpy(), **tol)
    np.testing.assert_allclose(gxa[:, frame], x_all.grad[:, frame].numpy(), **tol)
    np.testing.assert_allclose(gva[:, frame], v_all.grad[:, frame].numpy(), **tol)
    assert (gha[:, 0] == 7.0).all() and (gxa[:, 2] == 7.0).all()      # only the observed frame's slice is written
    if mode > 0:
        np.testing.assert_allclose(ghi, h_int.grad.numpy(), **tol)
        np.testing.assert_allclose(gxi, x_int.grad.numpy(), **tol)
        np.testing.assert_allclose(gvi, v_int.grad.numpy(), **tol)
    if mode == 2:
        ref = torch.cat([W0.grad.reshape(-1), b0.grad, w2.grad, b2.grad]).numpy() * 2.0    # written, then accumulated
        np.testing.assert_allclose(g_attn[:-1], ref[:-1], rtol=2e-4, atol=2e-5)
        assert abs(g_attn[-1]) < 1e-4      # softmax is shift invariant: dL/db2 = 0 up to rounding


def test_segno_embedding_of_all_frames_and_accumulate_kernels():
    """nb_segno_embed_forward / _backward over the rows of every observed frame (model.py:73) and nb_accumulate."""
    L = E.lib()
    g = torch.Generator().manual_seed(3)
    rows, F = 150, 1
    cfg = E.cabi.NbSegnoConfig(5, 5, 4, F, 2, 1, 1.0, 1)
    import ctypes
    total = int(L.nb_segno_param_count(ctypes.byref(cfg)))
    params = E.f32(torch.randn(total, generator=g) * 0.3)
    his = torch.randn(rows, F, generator=g)
    W, b = torch.tensor(params[:64 * F]).view(64, F).requires_grad_(True), torch.tensor(params[64 * F:64 * F + 64]).requires_grad_(True)
    ref = his @ W.T + b
    Gh = torch.randn(rows, 64, generator=g)
    (ref * Gh).sum().backward()
    h = np.zeros((rows, 64), np.float32)
    E.check(L.nb_segno_embed_forward(ctypes.byref(cfg), E.ptr(params), rows, E.ptr(E.f32(his)), E.ptr(h), None))
    np.testing.assert_allclose(h, ref.detach().numpy(), rtol=1e-6, atol=1e-6)
    gp = np.full(total, 3.0, np.float32)
    ws = np.zeros(int(L.nb_segno_embed_backward_workspace_floats(ctypes.byref(cfg), rows)), np.float32)
    E.check(L.nb_segno_embed_backward(ctypes.byref(cfg), rows, E.ptr(E.f32(his)), E.ptr(E.f32(Gh)), E.ptr(gp), E.ptr(ws), None))
    np.testing.assert_allclose(gp[:64 * F], W.grad.reshape(-1).numpy(), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(gp[64 * F:64 * F + 64], b.grad.numpy(), rtol=2e-5, atol=2e-5)
    assert (gp[64 * F + 64:] == 3.0).all()      # only the embedding entries are written
    a, c = E.f32(torch.randn(1000, generator=g)), E.f32(torch.randn(1000, generator=g))
    want = a + c
    E.check(L.nb_accumulate(1000, E.ptr(a), E.ptr(c), None))
    np.testing.assert_array_equal(a, want)
