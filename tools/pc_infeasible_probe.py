#!/usr/bin/env python
"""Collect PC instances the kernel reports INFEASIBLE (soak: 0.14 % of anymal_b stand states) for an offline check with the oracle."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g
g.build()
from quadruped_drake_b200.controller import BatchedController
from quadruped_drake_b200.synth import generate
ctl = BatchedController("anymal_b", device=0)
q, v, traj, contact = generate(ctl.model, 1 << 18, 777 + 13 * 5, "stand", ctl.fk)
out = ctl.step("pc", q, v, traj, contact, debug=True)
bad = np.nonzero(out.status == 2)[0][:24]
good = np.nonzero(out.status == 0)[0][:8]
sel = np.concatenate([bad, good])
np.savez(ROOT / "gpurun_out" / "pc_infeasible.npz", q=q[sel], v=v[sel], traj=traj[sel], contact=contact[sel], status=out.status[sel], tau=out.tau[sel],
         iters=out.qp_info[sel, 3], res=out.qp_info[sel, 1])
print("bad", len(bad), "of", (out.status == 2).sum())
