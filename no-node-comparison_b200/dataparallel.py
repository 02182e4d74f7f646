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


class PeerGradBucket:
    """The flat gradient buffer of a model in symmetric memory: every rank's buffer is mapped into every process
    (NVLink / NVSwitch peer access), so the data-parallel reduction needs no collective call — `FlatAdam` reads all
    ranks' gradients in place and sums them inside its update kernel (`nb_adam_step_peers`), between two device-side
    barriers.  PyTorch's symmetric-memory allocator is used for the plumbing only (allocation, handle exchange, barrier)."""

    def __init__(self, numel: int, device: torch.device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        self.group = group
        self.buf = symm.empty(numel, dtype=torch.float32, device=device)
        self.buf.zero_()
        try:
            self.hdl = symm.rendezvous(self.buf, group)
        except Exception:                                        # older releases want the group registered first
            symm.enable_symm_mem_for_group(group.group_name)
            self.hdl = symm.rendezvous(self.buf, group)
        self.world, self.rank = int(self.hdl.world_size), int(self.hdl.rank)
        if self.world > 8:
            raise ValueError("peer-memory gradient reduction supports at most 8 ranks (one NVSwitch box)")
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]       # rank r's buffer as seen from this process
        if self.ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric memory: the local buffer pointer does not match its own mapping")
        self.numel = numel

    def barrier(self) -> None:
        """All ranks' current streams meet here (device-side signal exchange; enqueue-only, CUDA-graph capturable)."""
        self.hdl.barrier()
