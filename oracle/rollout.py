"""ORACLE (test infrastructure only): CPU restatement of the closed-loop rollout of csrc/wbc_rollout.cuh.

The reference's loop is Drake's Simulator around a discrete MultibodyPlant (simulate.py:38,160-182); its contact solver
is not restatable (third party, absent), so - like the product - the oracle advances the state with the QP's own
contact-consistent accelerations. Parity is unpinned against Drake; the integrator is pinned by its own known answers
(tests/test_oracle_rollout.py): constant-twist motion, quaternion norm, exactness for linear motion."""
from __future__ import annotations

import numpy as np

from . import controllers as oc
from . import trajectory as tr


def integrate(q, v, vd, dt):
    """v+ = v + dt vd; q+ = q + dt N(q) v+ (world-frame angular velocity: qdot = 1/2 [0, w] (x) q), renormalised."""
    vn = v + dt * vd
    qw, qx, qy, qz = q[0:4]
    wx, wy, wz = vn[0:3]
    quat = np.array([qw + 0.5 * dt * (-wx * qx - wy * qy - wz * qz),
                     qx + 0.5 * dt * (wx * qw + wy * qz - wz * qy),
                     qy + 0.5 * dt * (wy * qw + wz * qx - wx * qz),
                     qz + 0.5 * dt * (wz * qw + wx * qy - wy * qx)])
    quat = quat * (1.0 / np.sqrt(quat @ quat))
    qn = q.copy()
    qn[0:4] = quat
    qn[4:7] = q[4:7] + dt * vn[3:6]
    qn[7:] = q[7:] + dt * vn[6:]
    return qn, vn


def rollout(robot, kind, plan, q, v, t, n_steps, dt, grid=None, wait_time=0.0):
    """One robot, n_steps steps. `grid`: stored timestamps for the planners/towr.py nearest-sample semantics, or None
    for continuous evaluation. Returns (q, v, t, tau_last, metrics per step)."""
    ctl = {"id": oc.IDController, "clf": oc.CLFController, "pc": oc.PCController, "mptc": oc.MPTCController}[kind](robot)
    q, v = np.array(q, float), np.array(v, float)
    log = []
    tau = None
    for _ in range(n_steps):
        if grid is not None:
            traj, contact, _ = tr.towr_planner_output(plan, grid, t, wait_time)
        else:
            traj, contact, _ = plan.sample(min(max(t, 0.0), plan.total_time()))
        ctl.V = ctl.err = ctl.res = ctl.Vdot = 0.0
        o = ctl.control_law(q, v, oc.traj_to_dict(traj, contact))
        tau = o.tau
        log.append(np.array(o.metrics, float))
        q, v = integrate(q, v, o.vd, dt)
        t = t + dt
    return q, v, t, tau, np.array(log)
