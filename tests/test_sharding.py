"""Multi-GPU host logic on CPU: instance partition + the optional gather, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quadruped_drake_b200.sharding import gather_rows, shard_range


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 4096, 10**7 + 3):
        for world in (1, 2, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    # stand-in for the per-rank controller output: row i carries its global instance index
    local = torch.arange(lo, hi, dtype=torch.float64)[:, None].repeat(1, 12)
    full = gather_rows(local, n, dst=0)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((full.numpy(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_over_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 101
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert full.shape == (n, 12) and np.array_equal(full[:, 0], np.arange(n)) and tmax == 2.0
