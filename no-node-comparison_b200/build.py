"""In-tree nvcc build of the C-ABI library for sm_100a (no JIT cache: the .so travels with the tree)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "nbody_b200.cu")
OUT = os.path.join(HERE, "libnbody_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-ldl"]


def _sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(ROOT, "include", "nbody_b200.h")]


def needs_build() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_library(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compile csrc/nbody_b200.cu -> libnbody_b200.so.  Returns the library path.
    `defines` / `out`: profiling builds (e.g. -DNB_STAGE_CLOCKS) written next to the product library."""
    if out is not None:
        force = True
    out = out or OUT
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: cannot build libnbody_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + ["-I", os.path.join(ROOT, "include"), SRC, "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
