"""Host-side harness for the CUDA emulator (TEST INFRASTRUCTURE ONLY).

Builds csrc/nbody_b200.cu with g++ against tests/emu/cuda_emu.h and calls the *same* C ABI with
numpy arrays, so kernel indexing / reductions can be checked against the oracle without a GPU.
The emulated library lives under tests/emu/_build and is never visible to the product package.
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
OUT = os.path.join(EMU_DIR, "_build", "libnbody_emu.so")
CSRC = os.path.join(ROOT, "no-node-comparison_b200", "csrc")


def _load_cabi():
    spec = importlib.util.spec_from_file_location("_nb_cabi", os.path.join(ROOT, "no-node-comparison_b200", "_cabi.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


cabi = _load_cabi()


def build_emu() -> str:
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(EMU_DIR, "cuda_emu.h"),
                                                                 os.path.join(EMU_DIR, "cuda_emu.cpp"),
                                                                 os.path.join(ROOT, "include", "nbody_b200.h")]
    if os.path.isfile(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-x", "c++", "-DNB_EMU", "-include", os.path.join(EMU_DIR, "cuda_emu.h"),
           "-I", os.path.join(ROOT, "include"), "-shared", "-fPIC", os.path.join(CSRC, "nbody_b200.cu"),
           os.path.join(EMU_DIR, "cuda_emu.cpp"), "-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + res.stderr)
    return OUT


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = cabi.declare(ctypes.CDLL(build_emu()))
    return _LIB


def ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().nb_last_error().decode())


def flat_params(weights: dict, order) -> np.ndarray:
    return np.concatenate([np.asarray(weights[k], dtype=np.float32).reshape(-1) for k in order])


def egno_run(cfg_kw: dict, params: np.ndarray, x, nodes, edge_fea, v, loc_mean, t_out, Gx=None, Gv=None, Gh=None, t_in=None):
    """forward (+ backward when cotangents are given) through the emulated C ABI."""
    L = lib()
    cfg = cabi.NbEgnoConfig(**cfg_kw)
    npar = L.nb_egno_param_count(ctypes.byref(cfg))
    assert npar == params.size, (npar, params.size)
    Nn = cfg.T * cfg.B * cfg.N
    x_out = np.zeros((Nn, 3), np.float32)
    v_out = np.zeros((Nn, 3), np.float32)
    h_out = np.zeros((Nn, 64), np.float32)
    saved = np.zeros(L.nb_egno_saved_floats(ctypes.byref(cfg)), np.float32)
    ws = np.zeros(L.nb_egno_workspace_floats(ctypes.byref(cfg), 0), np.float32)
    x, nodes, edge_fea, v, loc_mean = f32(x), f32(nodes), f32(edge_fea), f32(v), f32(loc_mean)
    t_out = np.ascontiguousarray(np.asarray(t_out, dtype=np.int64))
    t_in = None if t_in is None else np.ascontiguousarray(np.asarray(t_in, dtype=np.int64))
    check(L.nb_egno_forward(ctypes.byref(cfg), ptr(params), ptr(x), ptr(nodes), ptr(edge_fea), ptr(v), ptr(loc_mean),
                            ptr(t_out), ptr(t_in), ptr(x_out), ptr(v_out), ptr(h_out), ptr(saved), ptr(ws), None))
    res = dict(x_out=x_out, v_out=v_out, h_out=h_out)
    if Gx is not None:
        ws2 = np.zeros(L.nb_egno_workspace_floats(ctypes.byref(cfg), 1), np.float32)
        gp = np.full(npar, np.nan, np.float32)
        lead = (cfg.num_inputs,) if cfg.num_inputs > 1 else ()
        gx = np.zeros(lead + (cfg.B * cfg.N, 3), np.float32)
        gv = np.zeros(lead + (cfg.B * cfg.N, 3), np.float32)
        check(L.nb_egno_backward(ctypes.byref(cfg), ptr(params), ptr(nodes), ptr(edge_fea), ptr(loc_mean), ptr(t_out),
                                 ptr(t_in), ptr(saved), ptr(f32(Gx)), ptr(f32(Gv)), ptr(f32(Gh)), ptr(gp), ptr(gx), ptr(gv),
                                 ptr(ws2), None))
        res.update(grad_params=gp, gx_in=gx, gv_in=gv)
    return res


def segno_run(cfg_kw: dict, params: np.ndarray, his, x, v, edge_attr, Gx=None, Gh=None, Gv=None):
    L = lib()
    cfg = cabi.NbSegnoConfig(**cfg_kw)
    npar = L.nb_segno_param_count(ctypes.byref(cfg))
    assert npar == params.size, (npar, params.size)
    Nn = cfg.B * cfg.N
    x_out = np.zeros((Nn, 3), np.float32)
    v_out = np.zeros((Nn, 3), np.float32)
    h_out = np.zeros((Nn, 64), np.float32)
    saved = np.zeros(L.nb_segno_saved_floats(ctypes.byref(cfg)), np.float32)
    ws = np.zeros(L.nb_segno_workspace_floats(ctypes.byref(cfg), 0), np.float32)
    his, x, v, edge_attr = f32(his), f32(x), f32(v), f32(edge_attr)
    check(L.nb_segno_forward(ctypes.byref(cfg), ptr(params), ptr(his), ptr(x), ptr(v), ptr(edge_attr), ptr(x_out),
                             ptr(h_out), ptr(v_out), ptr(saved), ptr(ws), None))
    res = dict(x_out=x_out, v_out=v_out, h_out=h_out)
    if Gx is not None:
        ws2 = np.zeros(L.nb_segno_workspace_floats(ctypes.byref(cfg), 1), np.float32)
        gp = np.full(npar, np.nan, np.float32)
        gx = np.zeros((Nn, 3), np.float32)
        gv = np.zeros((Nn, 3), np.float32)
        check(L.nb_segno_backward(ctypes.byref(cfg), ptr(params), ptr(his), ptr(edge_attr), ptr(saved), ptr(f32(Gx)),
                                  ptr(f32(Gh)), ptr(f32(Gv)), ptr(gp), ptr(gx), ptr(gv), None, ptr(ws2), None))
        res.update(grad_params=gp, gx_in=gx, gv_in=gv)
    return res
