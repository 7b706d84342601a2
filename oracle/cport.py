"""ORACLE (test infrastructure / CPU baseline only): ctypes front end of oracle/c/oracle_id.c."""
from __future__ import annotations

import ctypes as C
import os
import time
from pathlib import Path

import numpy as np

LIB = Path(__file__).resolve().parent / "_build" / "liboracle_c.so"


def _lib():
    lib = C.CDLL(str(LIB))
    lib.oracle_id_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 8 + [C.c_int]
    lib.oracle_id_batch.restype = C.c_int
    return lib


def id_batch(robot, q, v, traj, contact, threads=None, **params):
    """IDController.ControlLaw for a batch on `threads` host threads -> tau, vd, f, status."""
    from quadruped_drake_b200 import load_robot
    from quadruped_drake_b200.capi import make_params
    lib = _lib()
    ms, pr = load_robot(robot).as_struct(), make_params(**params)
    q, v, traj = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, v, traj))
    contact = np.ascontiguousarray(contact, dtype=np.uint8)
    n = len(q)
    tau, vd, f, st = np.zeros((n, 12)), np.zeros((n, 18)), np.zeros((n, 4, 3)), np.zeros(n, np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    lib.oracle_id_batch(C.byref(ms), C.byref(pr), n, p(q), p(v), p(traj), p(contact), p(tau), p(vd), p(f), p(st),
                        threads or os.cpu_count() or 1)
    return tau, vd, f, st


def time_id_steps(robot, q, v, traj, contact, budget_s=20.0, single_thread_s=3.0, **_ignored):
    """Throughput of the C port over a bounded sample (bench.py cpu_baseline): all host cores, and one thread."""
    cores = os.cpu_count() or 1
    n0 = min(len(q), 8 * cores)
    t0 = time.perf_counter()
    id_batch(robot, q[:n0], v[:n0], traj[:n0], contact[:n0], cores)
    per = (time.perf_counter() - t0) / n0
    n = int(max(n0, min(len(q), budget_s / max(per, 1e-9))))
    # about 20 s of CPU work (all cores x ~1.3 s of wall time): several passes over the sample when one pass is shorter, median pass
    passes = int(max(1, min(16, round(min(budget_s, 20.0) / max(cores * per * n, 1e-9)))))
    walls = []
    for _ in range(passes):
        t0 = time.perf_counter()
        _, _, _, st = id_batch(robot, q[:n], v[:n], traj[:n], contact[:n], cores)
        walls.append(time.perf_counter() - t0)
    wall = float(np.median(walls))
    single = None
    if single_thread_s > 0:
        n1 = int(max(8, min(len(q), single_thread_s / max(per * cores, 1e-9))))
        t0 = time.perf_counter()
        id_batch(robot, q[:n1], v[:n1], traj[:n1], contact[:n1], 1)
        single = {"value": n1 / (time.perf_counter() - t0), "unit": "steps/s", "sample": f"{n1} instances on one thread"}
    return {"value": n / wall, "unit": "steps/s", "cores": cores, "kind": "port",
            "single_thread": single,
            "not_converged": int((st != 0).sum()), "build": "gcc -O3 -march=native",
            "label": "CPU restatement of the reference path (C port of the oracle), NOT Drake + OSQP",
            "passes": passes,
            "sample": f"{passes} passes (median) over {n} of the {len(q)} instances of one batch, C restatement of the reference path (oracle/c/oracle_id.c: "
                      f"18-pass mass matrix + full-size dense IPM QP), one instance stream per host thread"}


def fk(robot):
    """Forward-kinematics callable for synth.generate backed by the C port (reference arm of bench.py)."""
    from quadruped_drake_b200 import load_robot
    lib = C.CDLL(str(LIB))
    ms = load_robot(robot).as_struct()
    p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731

    def _fk(q, v):
        n = len(q)
        P, Vf = np.zeros((n, 4, 3)), np.zeros((n, 4, 3))
        M, Cv, tg, J, Jdv = np.zeros((18, 18)), np.zeros(18), np.zeros(18), np.zeros((4, 3, 18)), np.zeros((4, 3))
        for i in range(n):
            qi, vi = np.ascontiguousarray(q[i]), np.ascontiguousarray(v[i])
            lib.oracle_dynamics(C.byref(ms), p(qi), p(vi), p(M), p(Cv), p(tg), p(J), p(Jdv), p(P[i]))
            Vf[i] = J @ vi
        return P, Vf
    return _fk
