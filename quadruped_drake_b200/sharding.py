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


class MultiGpuController:
    """All GPUs of the box behind one call (north star: instances shard trivially across the 8 GPUs, no NCCL beyond an optional
    host gather). One process, one library handle per device (`wbc_multi_create`); `step` splits the batch into contiguous
    shards (`shard_range`), runs them side by side and returns host arrays holding every instance's result - the host gather.
    Page-locked buffers (`pinned(n)`) let the kernels read / write host memory directly; pageable arrays are staged."""

    def __init__(self, robot="mini_cheetah", devices=None, dof_order="depth_first", **params):
        import ctypes as C
        from . import capi
        from .model import RobotModel, load_robot
        self.lib = capi.load_library()
        self.model = robot if isinstance(robot, RobotModel) else load_robot(robot, dof_order=dof_order)
        self.params = capi.make_params(**params)
        if devices is None:
            import torch
            devices = list(range(torch.cuda.device_count()))
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise RuntimeError("MultiGpuController: no CUDA device (there is no CPU fallback)")
        self._m = C.c_void_p()
        ms = self.model.as_struct()
        dev = (C.c_int * len(self.devices))(*self.devices)
        rc = self.lib.wbc_multi_create(C.byref(ms), C.byref(self.params), len(self.devices), dev, C.byref(self._m))
        if rc != 0:
            msg = self.lib.wbc_multi_last_error(self._m).decode() if self._m else "allocation failed"
            self.close()
            raise RuntimeError(f"wbc_multi_create failed ({rc}): {msg}")

    def close(self):
        if getattr(self, "_m", None):
            self.lib.wbc_multi_destroy(self._m)
            import ctypes as C
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.wbc_multi_launch_count(self._m))

    @staticmethod
    def pinned(n):
        """Page-locked input / output arrays for n instances (dict: q v traj contact tau metrics status)."""
        import numpy as np
        from . import capi
        return {"q": capi.pinned_empty((n, 19)), "v": capi.pinned_empty((n, 18)), "traj": capi.pinned_empty((n, 54)),
                "contact": capi.pinned_empty((n, 4), np.uint8), "tau": capi.pinned_empty((n, 12)),
                "metrics": capi.pinned_empty((n, 4)), "status": capi.pinned_empty((n,), np.int32)}

    def step(self, kind, q, v, traj, contact, tau=None, metrics=None, status=None):
        """tau[N,12], metrics[N,4], status[N] for host arrays of N instances (outputs may be passed in, e.g. page-locked)."""
        import ctypes as C
        import numpy as np
        from . import capi
        from .controller import StepOutput
        k = capi.KINDS[kind] if isinstance(kind, str) else int(kind)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 19)
        n = q.shape[0]
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(n, 18)
        traj = np.ascontiguousarray(traj, dtype=np.float64).reshape(n, 54)
        contact = np.ascontiguousarray(contact, dtype=np.uint8).reshape(n, 4)
        tau = np.empty((n, 12)) if tau is None else tau
        metrics = np.empty((n, 4)) if metrics is None else metrics
        status = np.empty(n, dtype=np.int32) if status is None else status
        io = capi.WbcIO(capi.np_ptr(q), capi.np_ptr(v), capi.np_ptr(traj), capi.np_ptr(contact), capi.np_ptr(tau),
                        capi.np_ptr(metrics), capi.np_ptr(status))
        rc = self.lib.wbc_multi_step_host(self._m, k, n, C.byref(io))
        if rc != 0:
            raise RuntimeError(f"wbc_multi_step_host failed ({rc}): {self.lib.wbc_multi_last_error(self._m).decode()}")
        return StepOutput(tau, metrics, status)
