"""Trajectory simulators: the numpy restatement against trajectories produced by the reference itself
(tests/golden/sim_*.npz), the CUDA kernels through the host emulator against both, and — on a GPU — the kernels through
the Python API against the same golden trajectories."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import sim_oracle as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHARGED = ["sim_charged_n5", "sim_charged_n20"]
GRAVITY = ["sim_gravity_n5", "sim_gravity_n20"]
# float64 with a different summation order than numpy's, amplified by the dynamics over <= 2 000 steps
TOL = 1e-9


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", CHARGED)
def test_oracle_charged_reproduces_the_reference(name):
    d = _load(name)
    for k in range(d["loc0"].shape[0]):
        loc, vel = S.simulate_charged(d["loc0"][k], d["vel0"][k], d["charges"][k], int(d["T"]), int(d["sample_freq"]))
        np.testing.assert_allclose(loc, d["loc"][k], rtol=0, atol=1e-12)
        np.testing.assert_allclose(vel, d["vel"][k], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", GRAVITY)
def test_oracle_gravity_reproduces_the_reference(name):
    d = _load(name)
    for k in range(d["pos0"].shape[0]):
        pos, vel, force = S.simulate_gravity(d["pos0"][k], d["vel0"][k], d["mass"][k], int(d["T"]), int(d["sample_freq"]))
        np.testing.assert_allclose(pos, d["pos"][k], rtol=0, atol=1e-11)
        np.testing.assert_allclose(vel, d["vel"][k], rtol=0, atol=1e-11)
        np.testing.assert_allclose(force, d["force"][k], rtol=0, atol=1e-10)


def test_initial_condition_samplers_follow_the_reference_stream():
    """Same seed -> the charges / masses the reference drew (they are stored in the golden files)."""
    np.random.seed(43)
    for k in range(2):
        l0, v0, q = S.charged_initial_conditions(5, 2000 // 100 - 1)
        d = _load("sim_charged_n5")
        np.testing.assert_array_equal(q, d["charges"][k])
        np.testing.assert_array_equal(l0, d["loc0"][k])
        assert np.allclose(np.sqrt((v0 ** 2).sum(0)), 0.5)


@pytest.mark.parametrize("name", CHARGED)
def test_emulated_charged_kernel(name):
    from tests import emu_harness as E
    L, d = E.lib(), _load(name)
    B, _, N = d["loc0"].shape
    T, sf = int(d["T"]), int(d["sample_freq"])
    ns = T // sf - 1
    loc, vel = np.zeros((B, ns, 3, N)), np.zeros((B, ns, 3, N))
    c = ctypes.c_double
    E.check(L.nb_sim_charged(B, N, T, sf, c(1e-3), c(1.0), c(100.0), c(5.0), E.ptr(np.ascontiguousarray(d["loc0"])),
                             E.ptr(np.ascontiguousarray(d["vel0"])), E.ptr(np.ascontiguousarray(d["charges"].reshape(B, N))),
                             E.ptr(loc), E.ptr(vel), None))
    np.testing.assert_allclose(loc, d["loc"], rtol=0, atol=TOL)
    np.testing.assert_allclose(vel, d["vel"], rtol=0, atol=TOL)
    assert L.nb_sim_charged(B, 129, T, sf, c(1e-3), c(1.0), c(100.0), c(5.0), E.ptr(loc), E.ptr(loc), E.ptr(loc), E.ptr(loc),
                            E.ptr(vel), None) < 0


@pytest.mark.parametrize("name", GRAVITY)
def test_emulated_gravity_kernel(name):
    from tests import emu_harness as E
    L, d = E.lib(), _load(name)
    B, N, _ = d["pos0"].shape
    T, sf = int(d["T"]), int(d["sample_freq"])
    ns = T // sf
    pos, vel, force = (np.zeros((B, ns, N, 3)) for _ in range(3))
    c = ctypes.c_double
    E.check(L.nb_sim_gravity(B, N, T, sf, c(1e-3), c(1.0), c(0.1), E.ptr(np.ascontiguousarray(d["pos0"])),
                             E.ptr(np.ascontiguousarray(d["vel0"])), E.ptr(np.ascontiguousarray(d["mass"].reshape(B, N))),
                             E.ptr(pos), E.ptr(vel), E.ptr(force), None))
    np.testing.assert_allclose(pos, d["pos"], rtol=0, atol=TOL)
    np.testing.assert_allclose(vel, d["vel"], rtol=0, atol=TOL)
    np.testing.assert_allclose(force, d["force"], rtol=0, atol=1e-8)


def test_reflection_at_the_box_walls_matches_the_oracle():
    """An initial position outside the +-5 box is reflected in place before the integration starts (_clamp)."""
    from tests import emu_harness as E
    L = E.lib()
    rng = np.random.RandomState(0)
    N, T, sf = 6, 400, 100
    loc0 = rng.randn(1, 3, N) * 2.0
    loc0[0, 0, 1], loc0[0, 2, 4] = 5.7, -6.1
    vel0 = rng.randn(1, 3, N)
    q = rng.choice([-1.0, 1.0], size=(1, N))
    ref_l, ref_v = S.simulate_charged(loc0[0], vel0[0], q[0], T, sf)
    loc, vel = np.zeros((1, T // sf - 1, 3, N)), np.zeros((1, T // sf - 1, 3, N))
    c = ctypes.c_double
    E.check(L.nb_sim_charged(1, N, T, sf, c(1e-3), c(1.0), c(100.0), c(5.0), E.ptr(loc0), E.ptr(vel0), E.ptr(q), E.ptr(loc),
                             E.ptr(vel), None))
    np.testing.assert_allclose(loc[0], ref_l, rtol=0, atol=TOL)
    np.testing.assert_allclose(vel[0], ref_v, rtol=0, atol=TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CHARGED + GRAVITY)
def test_gpu_simulators_reproduce_the_reference_trajectories(name):
    import no_node_comparison_b200 as nb
    dev = torch.device("cuda:0")
    d = _load(name)
    T, sf = int(d["T"]), int(d["sample_freq"])
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    if "charged" in name:
        loc, vel = nb.simulate_charged(t(d["loc0"]), t(d["vel0"]), t(d["charges"]), T, sf)
        np.testing.assert_allclose(loc.cpu().numpy(), d["loc"], rtol=0, atol=TOL)
        np.testing.assert_allclose(vel.cpu().numpy(), d["vel"], rtol=0, atol=TOL)
    else:
        pos, vel, force = nb.simulate_gravity(t(d["pos0"]), t(d["vel0"]), t(d["mass"]), T, sf)
        np.testing.assert_allclose(pos.cpu().numpy(), d["pos"], rtol=0, atol=TOL)
        np.testing.assert_allclose(vel.cpu().numpy(), d["vel"], rtol=0, atol=TOL)
        np.testing.assert_allclose(force.cpu().numpy(), d["force"], rtol=0, atol=1e-8)
    with pytest.raises(ValueError):
        nb.simulate_charged(t(d.get("loc0", d.get("pos0"))).float(), None, None, T, sf)
