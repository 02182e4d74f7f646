#!/usr/bin/env python
"""Install the UNMODIFIED reference into ``oracle/_ref/`` (TEST INFRASTRUCTURE).

    python oracle/make_ref.py            # copies /root/reference/**/*.py (+ yaml) -> oracle/_ref/

The reference is pure Python (no setup.py / pyproject: nothing to ``pip install``), so "installing" it is a file copy.
``oracle/_ref/`` is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so the copy
travels to the GPU box, where ``/root/reference`` does not exist.  It is used only as the checker / baseline:

  * ``tests/test_driver_dropin.py``  – the reference's own ``run_epoch`` / ``rollout_fn`` driving its own modules on the
                                       CPU and the CUDA drop-in modules on the GPU, losses compared;
  * ``bench.py --impl reference``    – times the reference's real ``EGNO`` / ``SEGNO`` modules (``kind: "reference"``).

``__graft_entry__.build()`` runs this whenever ``/root/reference`` is present (the build container).
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("NB_REFERENCE_SRC", "/root/reference")
KEEP_EXT = (".py", ".yaml", ".yml", ".json", ".txt", ".md")
SKIP_DIRS = {".git", "__pycache__"}                    # data blobs (.pkl, .npy) are excluded by extension
MAX_BYTES = 512 * 1024                                   # no data blobs
# dead legacy entry points; EGNO_sweep.py carries a credential in a comment and is never copied
SKIP_FILES = {"EGNO_sweep.py", "sweep_params.py", "plotting.py", "testing.py", "artifact_model_map_complete.json"}


def install(src: str = SRC, dst: str = DST) -> int:
    if not os.path.isfile(os.path.join(src, "EGNO", "model", "egno.py")):
        raise RuntimeError(f"no reference under {src}")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        for f in files:
            p = os.path.join(root, f)
            if f in SKIP_FILES or not f.endswith(KEEP_EXT) or os.path.getsize(p) > MAX_BYTES:
                continue
            out = os.path.join(dst, os.path.relpath(p, src))
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(p, out)
            n += 1
    with open(os.path.join(dst, "INSTALLED_FROM"), "w") as fh:
        fh.write(f"{src}\n{n} files, unmodified\n")
    return n


if __name__ == "__main__":
    print(f"installed {install()} reference files into {DST}")
    sys.exit(0)
