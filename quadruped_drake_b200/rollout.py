"""Closed-loop batched rollout: the caller of the control step, reference simulate.py:160-182.

    planner (device sampler) -> controller step -> semi-implicit Euler, n_steps times for N robots, no host round trips.

Two plants: `plant=False` advances the state with the QP's own contact-consistent accelerations ("planned contacts hold",
csrc/wbc_rollout.cuh); `plant=True` applies the controller's TORQUES to the simulated robot on flat ground (forward dynamics +
velocity-level contact with Coulomb friction, csrc/wbc_plant.cuh), so that a controller can fail physically. Everything runs
through `wbc_rollout_ex` in the C ABI. There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import KINDS, WbcPlantOpts, WbcRolloutIO, WbcRolloutOpts, np_ptr
from .model import NQ, NU, NV

Q0_MINI_CHEETAH = np.array([1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.3] + [0.0, -0.8, 1.6] * 4)     # simulate.py:171-176


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class RolloutResult:
    __slots__ = ("q", "v", "t", "tau", "metrics", "status_or", "err_max", "metrics_log", "f_contact")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


def _opts(use_graph, plant, mu, erp, iters, f_ptr=None):
    return WbcRolloutOpts(1 if use_graph else 0, 1 if plant else 0, WbcPlantOpts(float(mu), float(erp), int(iters), 0), f_ptr)


def rollout(ctl, sampler, kind, q, v, t, n_steps, dt=5e-3, plan_index=None, log_metrics=False, use_graph=True, plant=False,
            mu=1.0, erp=0.2, iters=30) -> RolloutResult:
    """Advance N robots by n_steps control periods of length dt (simulate.py:20-21: dt = 5e-3, 6 s = 1200 steps).
    NumPy inputs are copied (host entry); torch CUDA tensors are updated IN PLACE on torch's current stream.
    plant=True: ground-contact simulation driven by the torques (mu: ground friction, simulate.py:44-46); the result then
    carries the last step's ground forces `f_contact[N,4,3]`."""
    k = KINDS[kind] if isinstance(kind, str) else int(kind)
    if _is_torch(q):
        import torch
        n, dev = q.shape[0], q.device
        mk = lambda shape, dt_=torch.float64: torch.empty(shape, dtype=dt_, device=dev)  # noqa: E731
        r = RolloutResult(q=q, v=v, t=t, tau=mk((n, NU)), metrics=mk((n, 4)), status_or=mk((n,), torch.int32), err_max=mk((n,)),
                          metrics_log=mk((n_steps, n, 4)) if log_metrics else None, f_contact=mk((n, 4, 3)) if plant else None)
        p = lambda x: None if x is None else C.c_void_p(x.data_ptr())  # noqa: E731
        io = WbcRolloutIO(p(q), p(v), p(t), p(plan_index), p(r.tau), p(r.metrics), p(r.status_or), p(r.err_max), p(r.metrics_log))
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        o = _opts(use_graph, plant, mu, erp, iters, p(r.f_contact))
        ctl._check(ctl.lib.wbc_rollout_ex(ctl._h, k, sampler._p, n, int(n_steps), float(dt), C.byref(io), C.byref(o), stream), "wbc_rollout_ex")
        return r
    q = np.array(q, dtype=np.float64).reshape(-1, NQ)
    n = len(q)
    v = np.array(v, dtype=np.float64).reshape(n, NV)
    t = np.array(np.broadcast_to(np.asarray(t, dtype=np.float64), (n,)))
    pi = None if plan_index is None else np.ascontiguousarray(plan_index, dtype=np.int32).reshape(n)
    r = RolloutResult(q=q, v=v, t=t, tau=np.empty((n, NU)), metrics=np.empty((n, 4)), status_or=np.empty(n, np.int32), err_max=np.empty(n),
                      metrics_log=np.empty((n_steps, n, 4)) if log_metrics else None, f_contact=np.zeros((n, 4, 3)) if plant else None)
    opt = lambda a: None if a is None else np_ptr(a)  # noqa: E731
    io = WbcRolloutIO(np_ptr(q), np_ptr(v), np_ptr(t), opt(pi), np_ptr(r.tau), np_ptr(r.metrics), np_ptr(r.status_or), np_ptr(r.err_max),
                      opt(r.metrics_log))
    o = _opts(use_graph, plant, mu, erp, iters, opt(r.f_contact))
    ctl._check(ctl.lib.wbc_rollout_ex_host(ctl._h, k, sampler._p, n, int(n_steps), float(dt), C.byref(io), C.byref(o)), "wbc_rollout_ex_host")
    return r


def plant_step(ctl, q, v, tau, dt=5e-3, mu=1.0, erp=0.2, iters=30, ctrl_status=None):
    """One time step of N simulated robots on flat ground (wbc_plant_step_host) for host arrays:
    -> q+[N,19], v+[N,18], f_contact[N,4,3] (ground forces LF RF LH RH, world axes), status[N]."""
    q = np.array(q, dtype=np.float64).reshape(-1, NQ)
    n = len(q)
    v = np.array(v, dtype=np.float64).reshape(n, NV)
    tau = np.ascontiguousarray(tau, dtype=np.float64).reshape(n, NU)
    f, st = np.zeros((n, 4, 3)), np.zeros(n, np.int32)
    cs = None if ctrl_status is None else np.ascontiguousarray(ctrl_status, dtype=np.int32).reshape(n)
    o = WbcPlantOpts(float(mu), float(erp), int(iters), 0)
    ctl._check(ctl.lib.wbc_plant_step_host(ctl._h, n, float(dt), C.byref(o), np_ptr(q), np_ptr(v), np_ptr(tau), None,
                                           None if cs is None else np_ptr(cs), np_ptr(st), np_ptr(f)), "wbc_plant_step_host")
    return q, v, f, st


def integrate(ctl, q, v, vd, dt, t=None):
    """One semi-implicit Euler step on torch CUDA tensors, in place (wbc_integrate)."""
    import torch
    p = lambda x: None if x is None else C.c_void_p(x.data_ptr())  # noqa: E731
    stream = C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)
    ctl._check(ctl.lib.wbc_integrate(ctl._h, q.shape[0], float(dt), p(q), p(v), p(vd), p(t), stream), "wbc_integrate")
