"""B200-native EGNO / SEGNO hot path (drop-in for the reference's torch.nn.Module classes).

    from no_node_comparison_b200 import EGNO, SEGNO

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/nbody_b200.h.  There is no CPU
fallback: importing is cheap, but constructing a model's forward without the compiled
`libnbody_b200.so` (see build.py) or without a CUDA device raises.
"""
from .egno import EGNO  # noqa: F401
from .segno import SEGNO  # noqa: F401
from .build import build_library  # noqa: F401
from ._lib import load_library, library_path  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .optim import FlatAdam  # noqa: F401
from .rollout import prepare_inputs, conserved_energy, egno_rollout, segno_rollout  # noqa: F401
from .loss import trajectory_mse  # noqa: F401
from .simulate import simulate_charged, simulate_gravity  # noqa: F401

__all__ = ["EGNO", "SEGNO", "GraphedStep", "FlatAdam", "prepare_inputs", "conserved_energy", "egno_rollout", "segno_rollout",
           "trajectory_mse", "simulate_charged", "simulate_gravity", "build_library", "load_library", "library_path"]
