"""Data-parallel helpers: trajectories are independent graphs, so a batch shards over ranks
with no data-path collective; training adds ONE all-reduce of the flat gradient bucket per step
(`model.enable_data_parallel()` -> functional._maybe_allreduce)."""
from __future__ import annotations

import os
from typing import Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous near-equal shard [lo, hi) of `n_items` trajectories for `rank`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str = "nccl"):
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, local_rank, world)."""
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def broadcast_parameters(model: torch.nn.Module, src: int = 0) -> None:
    """Make every rank start from rank `src`'s weights (one broadcast per parameter; setup only)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=src)
