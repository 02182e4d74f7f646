"""Loader of the compiled C-ABI library.  Fails loudly: there is no fallback path."""
from __future__ import annotations

import ctypes
import os

from . import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path() -> str:
    # NB_B200_LIBRARY: a profiling build of the same sources (tools/stage_clocks.py); never a different implementation
    return os.environ.get("NB_B200_LIBRARY") or os.path.join(_HERE, "libnbody_b200.so")


def load_library() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.isfile(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m no_node_comparison_b200.build` "
                "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
        _LIB = _cabi.declare(ctypes.CDLL(path))
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().nb_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg} (code {rc})")
