"""Data-parallel plumbing on CPU: world_size-2 `gloo` run of the flat-bucket gradient all-reduce and of the
batch sharding used by bench.py (the CUDA kernels are not involved; they are covered by -m gpu tests)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from no_node_comparison_b200.functional import _maybe_allreduce
    from no_node_comparison_b200.dataparallel import shard_range

    g = torch.arange(10, dtype=torch.float32) * (rank + 1)          # per-rank "flat gradient bucket"
    _maybe_allreduce(g, (dist.group.WORLD, True))                   # mean (any optimizer)
    s = torch.arange(10, dtype=torch.float32) * (rank + 1)
    _maybe_allreduce(s, (dist.group.WORLD, False))                  # sum (1 / world rides in FlatAdam's kernel)
    # parameters multiplied on the torch side (multi-input SEGNO): gradient reduced by dp_param
    from no_node_comparison_b200.functional import dp_param
    w = torch.ones(3, requires_grad=True)
    ((rank + 1.0) * dp_param(w, (dist.group.WORLD, True)).sum() + dp_param(w, None).sum() * 0.0).backward()
    assert torch.allclose(w.grad, torch.full((3,), 1.5)), w.grad
    lo, hi = shard_range(11, rank, world)
    out.put((rank, g.tolist(), (lo, hi), s.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = (torch.arange(10, dtype=torch.float32) * 1.5).tolist()   # mean of 1x and 2x
    assert res[0][1] == expect and res[1][1] == expect
    assert res[0][2] == (0, 6) and res[1][2] == (6, 11)               # contiguous, covering, near-equal shards
    total = (torch.arange(10, dtype=torch.float32) * 3.0).tolist()
    assert res[0][3] == total and res[1][3] == total
