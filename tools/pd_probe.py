import sys, numpy as np
from pathlib import Path; R = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(R)); sys.path.insert(0, str(R / 'tests'))
import __graft_entry__ as g; g.build()
from quadruped_drake_b200.controller import BatchedController
from quadruped_drake_b200 import planner as pl
from quadruped_drake_b200.rollout import rollout
from test_gpu_rollout import grounded
for kw in ({}, {"pd_kd": 0.3}, {"pd_kd": 0.3, "pd_kp": 60.0}):
    ctl = BatchedController("mini_cheetah", device=0, **kw)
    n = 64
    q0 = grounded(ctl, n, np.random.default_rng(1), 0.02)
    bh = float(grounded(ctl, 1)[0, 6])
    s2 = pl.TrajectorySampler(ctl, pl.make_motion_plan("mini_cheetah", "standing", 2.0, base_height=bh))
    for dt, steps in ((5e-3, 200), (1e-3, 1000), (5e-4, 2000)):
        r = rollout(ctl, s2, "pd", q0, np.zeros((n, 18)), np.zeros(n), steps, dt, plant=True)
        print(kw, dt, np.unique(r.status_or), r.q[:, 6].min(), r.q[:, 6].max(), np.abs(r.v).max(), r.f_contact[:, :, 2].sum(axis=1).mean())
