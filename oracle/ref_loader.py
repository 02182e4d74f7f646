"""Import the *real* reference modules (TEST INFRASTRUCTURE).

Where the reference is looked for, in order: ``$NB_REFERENCE_ROOT``, ``/root/reference`` (the build container), and
``oracle/_ref/`` — the unmodified copy that ``oracle/make_ref.py`` installs (git-ignored, travels to the GPU box).
Nothing under ``no-node-comparison_b200/`` imports this module.  Users: ``tests/golden/make_golden.py`` (fixture
generation), the ``not gpu`` cross-checks of the restatement, ``tests/test_driver_dropin.py`` (the reference's own
``run_epoch`` / ``rollout_fn`` driving the CUDA modules) and ``bench.py --impl reference`` (the CPU arm).

The reference needs ``torch_geometric`` and ``matplotlib`` only for three trivial symbols; they are
absent from this image, so tiny stand-ins are registered in ``sys.modules`` before importing.  Two shims make the
reference's drivers runnable at HEAD: ``EGNO.utils.random_ascending_tensor`` (imported from the wrong
module by ``EGNO/main_simulation_simple_no.py:8``) and — only on request, ``fix_segno_forward=True`` — the intended
``SEGNO.forward`` (HEAD's returns its inputs, ``SEGNO/models/model.py:92``).  No reference source is copied or edited.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    cands = [os.environ.get("NB_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "EGNO", "model", "egno.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_root()


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
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT} (run oracle/make_ref.py in the build container)")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # `import SEGNO` would execute SEGNO/__init__.py -> dataset_nbody -> torch_geometric;
    # pre-register a bare package object so only the model sub-package is imported.
    if "SEGNO" not in sys.modules:
        pkg = types.ModuleType("SEGNO")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "SEGNO")]
        sys.modules["SEGNO"] = pkg
    with contextlib.redirect_stdout(io.StringIO()):
        from EGNO.model.egno import EGNO  # type: ignore
        from SEGNO.models.model import SEGNO  # type: ignore
        import synthetic_sim  # type: ignore
    ns = types.SimpleNamespace(EGNO=EGNO, SEGNO=SEGNO, synthetic_sim=synthetic_sim, root=REFERENCE_ROOT)
    return ns


def segno_intended_forward(self, his, x, edges, v, edge_attr, T=10, in_steps=None):
    """The semantics every caller of SEGNO.forward assumes (SURVEY.md §0 defect 1): the integrated state of
    `forward_step` on the embedded input.  Used to drive the reference's own `train_nbody.run_epoch`, which cannot
    train with HEAD's literal forward (`model.py:92` returns the inputs)."""
    return self.forward_step(self.embedding(his), x, edges, v, edge_attr, T=T)


def load_reference_drivers():
    """The reference's callers of the hot path, unmodified: EGNO `run_epoch` / `rollout_fn` / `prepare_inputs`
    (EGNO/main_simulation_simple_no.py:190-384) with its dataset class, SEGNO `run_epoch` / `rollout_fn`
    (SEGNO/train_nbody.py:57-236) with its dataset class, and the metrics module `utils`."""
    ns = load_reference()
    import wandb  # the drivers call wandb.log unconditionally (main_simulation_simple_no.py:302, train_nbody.py:181)

    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("WANDB_SILENT", "true")
    if wandb.run is None:
        wandb.init(mode="disabled")
    with contextlib.redirect_stdout(io.StringIO()):
        import utils as ref_utils  # type: ignore
        import EGNO.utils as egno_utils  # type: ignore

        if not hasattr(egno_utils, "random_ascending_tensor"):     # SURVEY.md §0 defect 2
            egno_utils.random_ascending_tensor = ref_utils.random_ascending_tensor
        import EGNO.main_simulation_simple_no as egno_main  # type: ignore
        from EGNO.simulation.dataset_simple import NBodyDynamicsDataset  # type: ignore
        import SEGNO.train_nbody as segno_train  # type: ignore
        import SEGNO.dataset_nbody as segno_data  # type: ignore
    ns.utils = ref_utils
    ns.egno_main = egno_main
    ns.EgnoDataset = NBodyDynamicsDataset
    ns.segno_train = segno_train
    ns.SegnoDataset = segno_data.NBodyDataset
    return ns
