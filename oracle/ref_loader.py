"""Import the *real* reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing
under ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this module; it is
used by ``tests/golden/make_golden.py`` (fixture generation) and by the
``not gpu`` tests that cross-check the restatement when the reference is present.

The reference needs ``torch_geometric`` and ``matplotlib`` only for three trivial
symbols; they are absent from this image, so tiny stand-ins are
registered in ``sys.modules`` before importing.  No reference source is copied.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "EGNO", "model", "egno.py"))


def _install_stubs() -> None:
    import torch

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if "torch_geometric" not in sys.modules:
        try:
            import torch_geometric  # noqa: F401
        except Exception:
            tg = types.ModuleType("torch_geometric")
            tgu = types.ModuleType("torch_geometric.utils")
            tgd = types.ModuleType("torch_geometric.data")

            def to_dense_batch(x, batch):
                # graphs are equal-sized on this path -> a reshape + all-true mask
                nb = int(batch.max().item()) + 1
                out = x.reshape(nb, x.shape[0] // nb, *x.shape[1:])
                mask = torch.ones(out.shape[:2], dtype=torch.bool, device=x.device)
                return out, mask

            class Data(dict):
                @classmethod
                def from_dict(cls, d):
                    return cls(d)

            class DataLoader(torch.utils.data.DataLoader):
                pass

            tgu.to_dense_batch = to_dense_batch
            tgd.Data = Data
            tgd.DataLoader = DataLoader
            tg.utils = tgu
            tg.data = tgd
            sys.modules["torch_geometric"] = tg
            sys.modules["torch_geometric.utils"] = tgu
            sys.modules["torch_geometric.data"] = tgd


def load_reference():
    """Returns a namespace with the reference's EGNO, SEGNO classes and synthetic_sim."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # `import SEGNO` would execute SEGNO/__init__.py -> dataset_nbody -> torch_geometric;
    # pre-register a bare package object so only the model sub-package is imported.
    if "SEGNO" not in sys.modules:
        pkg = types.ModuleType("SEGNO")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "SEGNO")]
        sys.modules["SEGNO"] = pkg
    import io
    import contextlib

    with contextlib.redirect_stdout(io.StringIO()):
        from EGNO.model.egno import EGNO  # type: ignore
        from SEGNO.models.model import SEGNO  # type: ignore
        import synthetic_sim  # type: ignore
    ns = types.SimpleNamespace(EGNO=EGNO, SEGNO=SEGNO, synthetic_sim=synthetic_sim)
    return ns
