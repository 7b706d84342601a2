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


# ------------------------------------------------------------------------------------------ ground-contact plant
def plant_step(plant, q, v, tau, dt, mu=1.0, erp=0.2, iters=30):
    """One time step of the simulated robot on flat ground: the numpy restatement of csrc/wbc_plant.cuh (same scheme, same
    sweep order, same iteration count), built on oracle.dynamics.Plant with a dense M^-1. Defined by this repository - Drake's
    implicit contact solver (simulate.py:38) is not restatable; what pins it are the invariants of tests/test_oracle_rollout.py.

        v_free = v + dt M^-1 (B tau - Cv - tau_g);  u = J_c v_free + A p,  A = J_c M^-1 J_c'
        per foot (LF RF LH RH), normal row then x, y:  0 <= p_n _|_ u_n + bias >= 0,  |p_t| <= mu p_n, u_t -> 0   (projected GS)
        v+ = v_free + M^-1 J_c' p;  q+ = q + dt N(q) v+
    -> q+, v+, ground forces f[4,3] = p / dt."""
    q, v, tau = np.asarray(q, float), np.asarray(v, float), np.asarray(tau, float)
    M, Cv, tau_g, S = plant.calc_dynamics(q, v)
    quant = [plant.frame_position_quantities(q, v, f) for f in plant.foot_frames]
    J = np.vstack([x[1] for x in quant])                    # 12 x 18, rows 3 k + i
    pz = np.array([x[0][2] for x in quant])
    Minv = np.linalg.inv(M)
    v_free = v + dt * (Minv @ (S.T @ tau - Cv - tau_g))
    X = Minv @ J.T
    A = J @ X
    u = J @ v_free
    bias = np.zeros(12)
    for k in range(4):
        bias[3 * k + 2] = (pz[k] if pz[k] > 0.0 else erp * pz[k]) / dt
    p = np.zeros(12)
    for _ in range(iters):
        for k in range(4):
            r = 3 * k + 2
            new = max(0.0, p[r] - (u[r] + bias[r]) / A[r, r])
            u += A[:, r] * (new - p[r])
            p[r] = new
            for i in range(2):
                r = 3 * k + i
                lim = mu * p[3 * k + 2]
                new = min(max(p[r] - u[r] / A[r, r], -lim), lim)
                u += A[:, r] * (new - p[r])
                p[r] = new
    vn = v_free + X @ p
    qn, _ = integrate(q, vn, np.zeros(18), dt)
    return qn, vn, (p / dt).reshape(4, 3)
