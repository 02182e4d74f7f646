"""The oracle restatement must reproduce the reference's own outputs and gradients
(golden vectors made by tests/golden/make_golden.py from the real reference)."""
import pytest
import torch

from oracle import nbody_oracle as O
from tests.helpers import (EGNO_CASES, EGNO_MULTI_CASES, SEGNO_CASES, SEGNO_MULTI_CASES, load_case, rel_err,
                           egno_inputs_from_case, egno_multi_inputs_from_case, segno_inputs_from_case,
                           segno_multi_inputs_from_case)

TOL_OUT = 2e-6    # fp32 vs fp32, same formulas, different summation order
TOL_GRAD = 2e-5


@pytest.mark.parametrize("name", EGNO_CASES)
def test_egno_oracle_matches_reference(name):
    d, w, g = load_case(name)
    c = egno_inputs_from_case(d)
    p = {k: t.clone().requires_grad_(True) for k, t in w.items()}
    x = c["x"].clone().requires_grad_(True)
    v = c["v"].clone().requires_grad_(True)
    xo, vo, ho = O.egno_forward(p, x, c["nodes"], c["row"], c["col"], c["edge_attr"], v, c["loc_mean"], c["t_out"],
                                n_layers=c["L"], num_timesteps=c["T"])
    assert rel_err(xo, torch.tensor(d["x_out"])) < TOL_OUT
    assert rel_err(vo, torch.tensor(d["v_out"])) < TOL_OUT
    assert rel_err(ho, torch.tensor(d["h_out"])) < TOL_OUT
    loss = (xo * torch.tensor(d["Gx"])).sum() + (vo * torch.tensor(d["Gv"])).sum() + (ho * torch.tensor(d["Gh"])).sum()
    loss.backward()
    assert rel_err(x.grad, torch.tensor(d["gx_in"])) < TOL_GRAD
    assert rel_err(v.grad, torch.tensor(d["gv_in"])) < TOL_GRAD
    for k in g:
        got = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        assert rel_err(got, g[k]) < TOL_GRAD, k


@pytest.mark.parametrize("name", EGNO_MULTI_CASES)
def test_egno_multi_input_oracle_matches_reference(name):
    """num_inputs > 1 / per-trajectory output times: egno_forward_multi against the reference's own outputs."""
    d, w, g = load_case(name)
    c = egno_multi_inputs_from_case(d)
    p = {k: t.clone().requires_grad_(True) for k, t in w.items()}
    x = c["x"].clone().requires_grad_(True)
    v = c["v"].clone().requires_grad_(True)
    xo, vo, ho = O.egno_forward_multi(p, x, c["nodes"], c["row"], c["col"], c["edge_attr"], v, c["loc_mean"], c["t_in"],
                                      c["t_out"], n_layers=c["L"], num_timesteps=c["T"])
    assert rel_err(xo, torch.tensor(d["x_out"])) < TOL_OUT
    assert rel_err(vo, torch.tensor(d["v_out"])) < TOL_OUT
    assert rel_err(ho, torch.tensor(d["h_out"])) < TOL_OUT
    loss = (xo * torch.tensor(d["Gx"])).sum() + (vo * torch.tensor(d["Gv"])).sum() + (ho * torch.tensor(d["Gh"])).sum()
    loss.backward()
    assert rel_err(x.grad, torch.tensor(d["gx_in"])) < TOL_GRAD
    assert rel_err(v.grad, torch.tensor(d["gv_in"])) < TOL_GRAD
    for k in g:
        got = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        assert rel_err(got, g[k]) < TOL_GRAD, k


@pytest.mark.parametrize("name", SEGNO_CASES)
def test_segno_oracle_matches_reference(name):
    d, w, g = load_case(name)
    c = segno_inputs_from_case(d)
    p = {k: t.clone().requires_grad_(True) for k, t in w.items()}
    x = c["x"].clone().requires_grad_(True)
    v = c["v"].clone().requires_grad_(True)
    xo, ho, vo = O.segno_forward(p, c["his"], x, c["row"], c["col"], v, c["edge_attr"], c["T"])
    assert rel_err(xo, torch.tensor(d["x_out"])) < TOL_OUT
    assert rel_err(vo, torch.tensor(d["v_out"])) < TOL_OUT
    assert rel_err(ho, torch.tensor(d["h_out"])) < TOL_OUT
    loss = (xo * torch.tensor(d["Gx"])).sum() + (vo * torch.tensor(d["Gv"])).sum() + (ho * torch.tensor(d["Gh"])).sum()
    loss.backward()
    assert rel_err(x.grad, torch.tensor(d["gx_in"])) < TOL_GRAD
    assert rel_err(v.grad, torch.tensor(d["gv_in"])) < TOL_GRAD
    for k in g:
        got = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        assert rel_err(got, g[k]) < TOL_GRAD, k


@pytest.mark.parametrize("name", SEGNO_MULTI_CASES)
def test_segno_multi_input_oracle_matches_reference(name):
    """Several input frames with 'sum' / 'attn' merging (model.py:65-90,105-139): intended semantics, computed by
    make_golden.py with the reference's own forward_step / prepare_node_inputs."""
    d, w, g = load_case(name)
    c = segno_multi_inputs_from_case(d)
    p = {k: t.clone().requires_grad_(True) for k, t in w.items()}
    x = c["x"].clone().requires_grad_(True)
    v = c["v"].clone().requires_grad_(True)
    xo, ho, vo = O.segno_forward_multi(p, c["his"], x, c["row"], c["col"], v, c["edge_attr"], c["T"], c["in_steps"], c["agg"])
    assert rel_err(xo, torch.tensor(d["x_out"])) < TOL_OUT
    assert rel_err(vo, torch.tensor(d["v_out"])) < TOL_OUT
    assert rel_err(ho, torch.tensor(d["h_out"])) < TOL_OUT
    loss = (xo * torch.tensor(d["Gx"])).sum() + (vo * torch.tensor(d["Gv"])).sum() + (ho * torch.tensor(d["Gh"])).sum()
    loss.backward()
    assert rel_err(x.grad, torch.tensor(d["gx_in"])) < TOL_GRAD
    assert rel_err(v.grad, torch.tensor(d["gv_in"])) < TOL_GRAD
    for k in g:
        if k == "enc_attn_net.attn_mlp.2.bias":   # softmax is shift invariant: this gradient is rounding noise (~1e-6)
            continue
        got = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        assert rel_err(got, g[k]) < TOL_GRAD, k
