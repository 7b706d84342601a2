#!/usr/bin/env python
"""Soak run of the control step: 2^20 random instances per (robot, controller, contact pattern, torque box) combination; counts
the status words, the iteration tail, and certifies a 32768-instance sample of every ID / CLF combination through the KKT
conditions of the reference QP (tests/kkt.py). Prints one JSON line per combination (-> profiles/r2_soak.jsonl)."""
import json
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
import kkt  # noqa: E402
from quadruped_drake_b200.controller import BatchedController  # noqa: E402
from quadruped_drake_b200.synth import generate  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
combos = [(r, k, p, tl) for r in ("mini_cheetah", "anymal_b") for k in ("id", "clf", "pc") for p in ("stand", "walk", "trot", "mixed") for tl in (0, 1)
          if not (k == "pc" and tl) and not (k == "clf" and tl and p != "walk")]
for robot, kind, pattern, tl in combos:
    params = {"torque_limits": 1} if tl else {}
    ctl = BatchedController(robot, device=0, **params)
    st_all, it_all, worst = {}, [], {"stationarity": 0.0, "comp": 0.0, "eq": 0.0, "ineq": 0.0, "dual": 0.0}
    for c in range(0, N, 1 << 18):
        n = min(1 << 18, N - c)
        q, v, traj, contact = generate(ctl.model, n, 777 + c // 1000 + 13 * len(pattern) + tl, pattern, ctl.fk)
        out = ctl.step(kind, q, v, traj, contact, debug=True)
        s, cnt = np.unique(out.status, return_counts=True)
        for a, b in zip(s, cnt):
            st_all[int(a)] = st_all.get(int(a), 0) + int(b)
        it_all.append(out.qp_info[:, 3])
        bad = out.status != 0
        assert (out.tau[bad] == 0).all()
        if kind in ("id", "clf") and c == 0:
            sel = np.nonzero(~bad)[0][:32768]
            o = SimpleNamespace(tau=out.tau[sel], vd=out.vd[sel], f=out.f[sel], qp_info=out.qp_info[sel], lam=out.lam[sel])
            cert = kkt.certificate(kind, ctl.dynamics(q[sel], v[sel]), ctl.model, q[sel], v[sel], traj[sel], contact[sel], o, params)
            for key in ("stationarity", "comp", "eq", "ineq"):
                worst[key] = max(worst[key], float(cert[key].max()))
            worst["dual"] = float(cert["dual"].min())
    it = np.concatenate(it_all)
    print(json.dumps({"robot": robot, "controller": kind, "pattern": pattern, "torque_limits": tl, "instances": N, "status": st_all,
                      "iterations": {"mean": float(it.mean()), "p999": float(np.percentile(it, 99.9)), "max": float(it.max())},
                      "kkt_worst_of_32768": worst if kind in ("id", "clf") else None}), flush=True)
    ctl.close()
