#!/usr/bin/env python
"""Active-set iteration histogram of the BASELINE workloads (GPU): mean / max / distribution of Goldfarb-Idnani pivots per
instance, size of the final active set, status counts. Writes one JSON object per workload to stdout (-> profiles/)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

g.build()
from quadruped_drake_b200.controller import BatchedController  # noqa: E402
from quadruped_drake_b200.synth import generate  # noqa: E402

WORK = [("mini_cheetah", "id", "stand", 4096, 20260119, {}),
        ("anymal_b", "id", "trot", 16384, 20260120, {"torque_limits": 1}),
        ("mini_cheetah", "clf", "walk", 65536, 20260121, {}),
        ("mini_cheetah", "pc", "walk", 65536, 20260121, {}),
        ("mini_cheetah", "id", "walk", 16384, 21, {"torque_limits": 1})]
for robot, kind, pattern, n, seed, params in WORK:
    ctl = BatchedController(robot, device=0, **params)
    q, v, traj, contact = generate(ctl.model, n, seed, pattern, ctl.fk)
    out = ctl.step(kind, q, v, traj, contact, debug=True)
    it = out.qp_info[:, 3].astype(int)
    nact = (out.lam > 0).sum(axis=1)
    st, cnt = np.unique(out.status, return_counts=True)
    rec = {"robot": robot, "controller": kind, "pattern": pattern, "n": n, "params": params,
           "iterations": {"mean": float(it.mean()), "max": int(it.max()), "p50": float(np.median(it)), "p99": float(np.percentile(it, 99)),
                          "histogram": np.bincount(it, minlength=1).tolist()},
           "final_active_set": {"mean": float(nact.mean()), "max": int(nact.max()), "histogram": np.bincount(nact).tolist()},
           "drops_mean": float((it - nact).mean()),
           "status": {int(s): int(c) for s, c in zip(st, cnt)}}
    print(json.dumps(rec))
    bad = out.status != 0
    if bad.any():
        np.savez(ROOT / "gpurun_out" / f"failed_{robot}_{kind}_{pattern}.npz", q=q[bad][:64], v=v[bad][:64], traj=traj[bad][:64],
                 contact=contact[bad][:64], status=out.status[bad][:64], iters=it[bad][:64])
    ctl.close()
