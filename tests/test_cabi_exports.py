"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/nbody_b200.h declares.
(No compute calls: there is no GPU in the CPU test tier.)"""
import ctypes
import os
import re

from tests.helpers import ROOT


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import no_node_comparison_b200 as nb

    path = nb.build_library()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_bindings_cover_the_header_and_queries_work():
    import no_node_comparison_b200 as nb
    from no_node_comparison_b200 import _cabi

    assert sorted(_cabi.EXPORTS) == _declared_functions()
    lib = nb.load_library()
    assert lib.nb_version() >= 1
    cfg = _cabi.NbEgnoConfig(256, 20, 10, 4, 2, 2, 2, 32, 1)
    assert lib.nb_egno_param_count(ctypes.byref(cfg)) == 201736      # SURVEY.md §8a1 [probed]
    scfg = _cabi.NbSegnoConfig(256, 20, 10, 1, 2, 1, 1.0)
    assert lib.nb_segno_param_count(ctypes.byref(scfg)) == 33602     # SURVEY.md §8a9 [probed]
    assert lib.nb_egno_saved_floats(ctypes.byref(cfg)) > 0
    assert lib.nb_egno_workspace_floats(ctypes.byref(cfg), 1) > 0
    bad = _cabi.NbEgnoConfig(4, 20, 10, 4, 7, 2, 2, 32, 1)           # modes > T//2+1: invalid, like the reference
    assert lib.nb_egno_param_count(ctypes.byref(bad)) < 0
    assert b"num_modes" in lib.nb_last_error()


def test_sass_is_sm100a():
    """The shipped cubin targets sm_100a (no PTX JIT, no other arch)."""
    import subprocess
    import no_node_comparison_b200 as nb

    out = subprocess.run(["cuobjdump", "-lelf", nb.build_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
