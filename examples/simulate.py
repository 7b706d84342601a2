#!/usr/bin/env python
"""The reference's simulate.py loop (simulate.py:20-22,106-139,160-182) with this repository's drop-in pieces and no Drake:

    BasicTrunkPlanner / TowrTrunkPlanner  ->  IDController | CLFController | PCController | MPTCController (LeafSystem mirror)
                                          ->  ground-contact plant step (wbc_plant_step)  ->  next state

One robot, dt = 5e-3, sim_time = 6.0 like the reference (BASELINE configs[0]); every control step goes through the same
`DoSetControlTorques` callback Drake would call. Needs a CUDA device (there is no CPU fallback).

    python examples/simulate.py [--controller id|clf|pc|mptc] [--planner standing|orientation|raise_foot|towr] [--sim-time 6.0]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def run(controller="id", planner="standing", sim_time=6.0, dt=5e-3, verbose=True):
    from quadruped_drake_b200 import controller as ctl_mod, planner as pl
    from quadruped_drake_b200.rollout import plant_step
    cls = {"id": ctl_mod.IDController, "clf": ctl_mod.CLFController, "pc": ctl_mod.PCController, "mptc": ctl_mod.MPTCController}[controller]
    c = cls("mini_cheetah", dt)                                   # reference: IDController(plant, dt, use_lcm=use_lcm)
    batched = c.batched
    if planner == "towr":
        p = pl.TowrTrunkPlanner(batched, robot="mini_cheetah")
        trunk = p.SetTrunkOutputs
    else:
        p = pl.BasicTrunkPlanner(robot="mini_cheetah")
        def trunk(t):
            {"standing": p.SimpleStanding, "orientation": lambda: p.OrientationTest(t), "raise_foot": lambda: p.RaiseFoot(t)}[planner]()
            return p.output_dict
    # initial state of simulate.py:171-179, lowered so that the feet touch the ground
    q = np.array([1.0, 0, 0, 0, 0, 0, 0.3] + [0.0, -0.8, 1.6] * 4)
    v = np.zeros(18)
    q[6] -= batched.dynamics(q[None], v[None])["p_feet"][0, :, 2].min()
    context = c.CreateDefaultContext()
    log, t0 = [], time.perf_counter()
    for k in range(int(round(sim_time / dt))):
        t = k * dt
        context.FixValue(0, np.hstack([q, v]))                    # quad_state port (simulate.py:121-127)
        context.FixValue(1, trunk(t))                             # trunk_input port
        tau = c.EvalOutput(context, 0)                            # quad_torques: DoSetControlTorques -> ControlLaw
        met = c.EvalOutput(context, 1)                            # output_metrics (simulate.py:142 logs them)
        qn, vn, f, st = plant_step(batched, q[None], v[None], tau[None], dt)
        if st[0] != 0 or c.last_status != 0:
            print(f"robot fell / controller failed at t = {t:.3f} s (status {int(st[0])}); the synthetic gait plans of make_gait_plan are "
                  "not dynamically consistent - a solved TOWR plan is needed for walking")
            break
        q, v = qn[0], vn[0]
        log.append([t, q[6], met[1], f[0, :, 2].sum()])
    wall = time.perf_counter() - t0
    run.last_wall = wall                                          # the loop alone, without the construction of the controller
    log = np.array(log)
    if verbose:
        print(f"{controller} / {planner}: {len(log)} steps in {wall:.2f} s wall ({len(log) / wall:.0f} control steps/s, real-time factor "
              f"{sim_time / wall:.1f}); final base height {q[6]:.4f} m, tracking error metric {log[-1, 2]:.2e}, ground force {log[-1, 3]:.2f} N")
    return q, v, log


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--controller", default="id", choices=["id", "clf", "pc", "mptc"])
    ap.add_argument("--planner", default="standing", choices=["standing", "orientation", "raise_foot", "towr"])
    ap.add_argument("--sim-time", type=float, default=6.0)
    a = ap.parse_args()
    run(a.controller, a.planner, a.sim_time)
