#!/usr/bin/env python
"""Throughput of the one-process multi-GPU call (wbc_multi_step_host / MultiGpuController): N host instances in page-locked
arrays -> N torques in host arrays, all visible GPUs side by side. One JSON line per batch size."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

g.build()
import torch  # noqa: E402
from quadruped_drake_b200.controller import BatchedController  # noqa: E402
from quadruped_drake_b200.sharding import MultiGpuController  # noqa: E402
from quadruped_drake_b200.synth import generate  # noqa: E402

ndev = torch.cuda.device_count()
one = BatchedController("mini_cheetah", device=0)
for n in (4096 * ndev, 131072 * ndev, 1 << 20, 4 << 20):
    parts = [generate(one.model, min(1 << 20, n - o), 20260119 + o, "stand", one.fk) for o in range(0, n, 1 << 20)]
    q, v, traj, contact = (np.concatenate([p[i] for p in parts]) for i in range(4))
    res = {}
    for devs in ([0], list(range(ndev))):
        multi = MultiGpuController("mini_cheetah", devices=devs)
        buf = multi.pinned(n)
        buf["q"][:], buf["v"][:], buf["traj"][:], buf["contact"][:] = q, v, traj, contact
        for _ in range(3):
            multi.step("id", buf["q"], buf["v"], buf["traj"], buf["contact"], buf["tau"], buf["metrics"], buf["status"])
        reps = 20 if n <= (1 << 20) else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            multi.step("id", buf["q"], buf["v"], buf["traj"], buf["contact"], buf["tau"], buf["metrics"], buf["status"])
        dt = (time.perf_counter() - t0) / reps
        assert (buf["status"] == 0).all()
        res[len(devs)] = n / dt
        if len(devs) == 1:
            ref = buf["tau"].copy()
        else:
            assert np.array_equal(ref, buf["tau"])
        multi.close()
        del buf
    print(json.dumps({"metric": "whole-body QP control steps/sec, host arrays in / host arrays out, one process", "instances": n, "gpus": ndev,
                      "steps_per_s_1gpu": res[1], "steps_per_s_all_gpus": res[ndev], "speedup": res[ndev] / res[1]}), flush=True)
