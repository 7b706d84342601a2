"""Instance sharding across GPUs: contiguous equal chunks of [0, N), one process per GPU, no
collective on the data path. The optional gather of the torque shards onto rank 0 is the only communication."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int):
    """Half-open instance range of `rank`; sizes differ by at most one, earlier ranks take the remainder."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, n_total: int, dst: int = 0):
    """Optional: collect the per-rank row blocks (e.g. tau[N_local, 12]) on `dst` with torch.distributed.
    Works with any backend (NCCL on GPUs, gloo on CPU). Returns the full [n_total, ...] tensor on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
