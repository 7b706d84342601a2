"""ORACLE (test infrastructure only): restatement of the reference ControlLaw methods.

Follows, block by block, reference controllers/inverse_dynamics_controller.py:103-234,
controllers/clf_controller.py:48-234 and controllers/pc_controller.py:43-255 (+
mptc_controller.py:30-57), with MultibodyPlant replaced by oracle.dynamics.Plant and
`OsqpSolver().Solve(mp)` replaced by oracle.qp.solve_qp on the *full-size* program the
reference builds (variables [vd(18); tau(12); f_j(3 each); (delta)]). **parity unpinned**
(no golden data in the reference; see oracle/dynamics.py header).

Declared tie-break: the reference cost does not touch tau or f, so
with >= 2 stance feet its optimum is a set. Both this oracle and the CUDA path add
    reg_f/2 |f|^2 + reg_tau/2 |tau|^2 + reg_vd/2 |vd|^2 (+ reg_f/2 delta^2 for PC's free slack)
to the reference objective; defaults reg_f = 1e-6, others 0.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_continuous_are

from .dynamics import Plant, rpy_from_matrix, rpy_rate_matrix
from .qp import solve_qp

FEET = ["lf", "rf", "lh", "rh"]

DEFAULTS = dict(
    id_kp_body_p=500.0, id_kd_body_p=50.0, id_kp_body_rpy=500.0, id_kd_body_rpy=50.0,
    id_kp_foot=100.0, id_kd_foot=20.0, id_w_body=10.0, id_w_foot=1.0,
    clf_q_body_p=5000.0, clf_q_body_pd=200.0, clf_q_body_rpy=5000.0, clf_q_body_rpyd=200.0,
    clf_q_foot_p=200.0, clf_q_foot_pd=20.0, clf_r=1.0, clf_w_delta=1000.0,
    pc_kp_body_p=100.0, pc_kd_body_p=10.0, pc_kp_body_rpy=100.0, pc_kd_body_rpy=10.0,
    pc_kp_foot=200.0, pc_kd_foot=20.0, pc_w_body=10.0, pc_w_foot=1.0,
    mu=0.7, contact_damping=100.0, reg_f=1e-6, reg_tau=0.0, reg_vd=0.0, torque_limits=0,
)


def traj_to_dict(traj, contact):
    """traj[54] (+ contact[4]) in the include/wbc.h layout -> the reference trunk dict
    (planners/simple.py:45-85)."""
    t = np.asarray(traj, float)
    d = {"p_body": t[0:3], "pd_body": t[3:6], "pdd_body": t[6:9],
         "rpy_body": t[9:12], "rpyd_body": t[12:15], "rpydd_body": t[15:18]}
    for i, f in enumerate(FEET):
        d["p_" + f] = t[18 + 3 * i:21 + 3 * i]
        d["pd_" + f] = t[30 + 3 * i:33 + 3 * i]
        d["pdd_" + f] = t[42 + 3 * i:45 + 3 * i]
    d["contact_states"] = [bool(c) for c in contact]
    d["f_cj"] = np.zeros((3, 4))
    d["u2_max"] = 0.0
    return d


def dict_to_traj(d):
    t = np.zeros(54)
    for k, key in enumerate(["p_body", "pd_body", "pdd_body", "rpy_body", "rpyd_body", "rpydd_body"]):
        t[3 * k:3 * k + 3] = d[key]
    for i, f in enumerate(FEET):
        t[18 + 3 * i:21 + 3 * i] = d["p_" + f]
        t[30 + 3 * i:33 + 3 * i] = d["pd_" + f]
        t[42 + 3 * i:45 + 3 * i] = d["pdd_" + f]
    return t, np.array([1 if c else 0 for c in d["contact_states"]], dtype=np.uint8)


def standing_dict(robot="mini_cheetah"):
    """BasicTrunkPlanner.SimpleStanding (planners/simple.py:39-85)."""
    if robot == "anymal_b":
        feet = [[0.34, 0.19, 0.0], [0.34, -0.19, 0.0], [-0.34, 0.19, 0.0], [-0.34, -0.19, 0.0]]
        pb = [0.0, 0.0, 0.5]
    else:
        feet = [[0.175, 0.11, 0.0], [0.175, -0.11, 0.0], [-0.2, 0.11, 0.0], [-0.2, -0.11, 0.0]]
        pb = [0.0, 0.0, 0.3]
    d = {}
    for f, p in zip(FEET, feet):
        d["p_" + f], d["pd_" + f], d["pdd_" + f] = np.array(p), np.zeros(3), np.zeros(3)
    d["contact_states"] = [True] * 4
    d["f_cj"] = np.zeros((3, 4))
    d["rpy_body"], d["p_body"] = np.zeros(3), np.array(pb)
    for k in ["rpyd_body", "pd_body", "rpydd_body", "pdd_body"]:
        d[k] = np.zeros(3)
    d["u2_max"] = 0.0
    return d


class StepResult:
    pass


class OracleController:
    def __init__(self, robot="mini_cheetah", dof_order="depth_first", **params):
        self.plant = robot if isinstance(robot, Plant) else Plant(robot, dof_order)
        self.p = dict(DEFAULTS)
        self.p.update(params)
        self.V = self.err = self.res = self.Vdot = 0.0

    # --- shared pieces of every ControlLaw ------------------------------------------------
    def _common(self, q, v, trunk):
        pl = self.plant
        c = StepResult()
        c.M, c.Cv, c.tau_g, c.S = pl.calc_dynamics(q, v)
        c.contact = [bool(x) for x in trunk["contact_states"]]
        c.swing = [not x for x in c.contact]
        c.p_feet_nom = np.array([trunk["p_" + f] for f in FEET], float)
        c.pd_feet_nom = np.array([trunk["pd_" + f] for f in FEET], float)
        c.pdd_feet_nom = np.array([trunk["pdd_" + f] for f in FEET], float)
        (R, c.p_body), c.J_body, c.Jdv_body = pl.frame_pose_quantities(q, v, pl.base_link)
        c.pd_body = (c.J_body @ v)[3:]
        c.rpy = rpy_from_matrix(R)
        c.N = rpy_rate_matrix(c.rpy)
        c.omega = (c.J_body @ v)[:3]
        c.rpyd = np.linalg.solve(c.N, c.omega)  # CalcRpyDtFromAngularVelocityInParent
        quant = [pl.frame_position_quantities(q, v, f) for f in pl.foot_frames]
        c.p_feet = np.array([x[0] for x in quant])
        c.J_feet = np.array([x[1] for x in quant])
        c.Jdv_feet = np.array([x[2] for x in quant])
        c.pd_feet = c.J_feet @ v
        return c

    def _constraints(self, c, v, n, extra_cols=0):
        """Dynamics, friction-pyramid and contact rows (inverse_dynamics_controller.py:48-101)
        for x = [vd(18), tau(12), f(3 nc), extra]."""
        nv, nu = 18, 12
        cont = [i for i in range(4) if c.contact[i]]
        nc = len(cont)
        A = np.zeros((nv + 3 * nc, n))
        b = np.zeros(nv + 3 * nc)
        A[:nv, :nv] = c.M
        A[:nv, nv:nv + nu] = -c.S.T
        for j, i in enumerate(cont):
            A[:nv, nv + nu + 3 * j:nv + nu + 3 * j + 3] = -c.J_feet[i].T
        b[:nv] = -c.Cv - c.tau_g
        G = np.zeros((4 * nc, n))
        mu = self.p["mu"]
        A_i = np.array([[1, 0, -mu], [-1, 0, -mu], [0, 1, -mu], [0, -1, -mu]], float)
        for j, i in enumerate(cont):
            G[4 * j:4 * j + 4, nv + nu + 3 * j:nv + nu + 3 * j + 3] = A_i
            A[nv + 3 * j:nv + 3 * j + 3, :nv] = c.J_feet[i]
            pd = c.J_feet[i] @ v
            b[nv + 3 * j:nv + 3 * j + 3] = -self.p["contact_damping"] * pd - c.Jdv_feet[i]
        h = np.zeros(4 * nc)
        if self.p["torque_limits"]:
            T = np.zeros((2 * nu, n))
            T[:nu, nv:nv + nu] = np.eye(nu)
            T[nu:, nv:nv + nu] = -np.eye(nu)
            G = np.vstack([G, T])
            h = np.hstack([h, self.plant.effort, self.plant.effort])
        return A, b, G, h, cont

    def _regularise(self, P, nc, delta_col=None, delta_reg=0.0):
        nv, nu = 18, 12
        P[:nv, :nv] += self.p["reg_vd"] * np.eye(nv)
        P[nv:nv + nu, nv:nv + nu] += self.p["reg_tau"] * np.eye(nu)
        k = nv + nu
        P[k:k + 3 * nc, k:k + 3 * nc] += self.p["reg_f"] * np.eye(3 * nc)
        if delta_col is not None:
            P[delta_col, delta_col] += delta_reg

    def _finish(self, c, res, P0, q0, A, b, G, h, cont, P, qv):
        out = StepResult()
        x = res.x
        out.vd, out.tau = x[:18], x[18:30]
        out.f = np.zeros((4, 3))
        for j, i in enumerate(cont):
            out.f[i] = x[30 + 3 * j:33 + 3 * j]
        out.x, out.nu, out.lam, out.active = x, res.nu, res.lam, res.active
        out.objective = 0.5 * x @ P0 @ x + q0 @ x          # the function the reference hands the solver
        out.objective_reg = 0.5 * x @ P @ x + qv @ x
        out.primal_res = max(np.abs(A @ x - b).max(), max(0.0, (G @ x - h).max()) if G.shape[0] else 0.0)
        out.status = res.status
        out.qp = (P, qv, A, b, G, h)
        out.common = c
        return out

    def metrics(self):
        """SetLoggingOutputs (basic_controller.py:271-283)."""
        return np.array([self.V, self.err, self.res, self.Vdot])


class IDController(OracleController):
    """inverse_dynamics_controller.py:103-234."""

    def control_law(self, q, v, trunk):
        p = self.p
        q, v = np.asarray(q, float), np.asarray(v, float)
        c = self._common(q, v, trunk)
        sw = [i for i in range(4) if c.swing[i]]
        pdd_body_des = trunk["pdd_body"] - p["id_kp_body_p"] * (c.p_body - trunk["p_body"]) \
            - p["id_kd_body_p"] * (c.pd_body - trunk["pd_body"])
        rpydd_des = trunk["rpydd_body"] - p["id_kp_body_rpy"] * (c.rpy - trunk["rpy_body"]) \
            - p["id_kd_body_rpy"] * (c.rpyd - trunk["rpyd_body"])
        omegad_des = c.N @ rpydd_des                     # no Ndot term
        vd_body_des = np.hstack([omegad_des, pdd_body_des])
        nc = sum(c.contact)
        n = 30 + 3 * nc
        P0, q0 = np.zeros((n, n)), np.zeros(n)

        def jac_cost(J, Jdv, xdd_des, w):               # AddJacobianTypeCost :25-35
            P0[:18, :18] += w * J.T @ J
            q0[:18] += w * (J.T @ (Jdv - xdd_des))
        jac_cost(c.J_body, c.Jdv_body, vd_body_des, p["id_w_body"])
        for i in sw:
            pdd_s_des = c.pdd_feet_nom[i] - p["id_kp_foot"] * (c.p_feet[i] - c.p_feet_nom[i]) \
                - p["id_kd_foot"] * (c.pd_feet[i] - c.pd_feet_nom[i])
            jac_cost(c.J_feet[i], c.Jdv_feet[i], pdd_s_des, p["id_w_foot"])
        A, b, G, h, cont = self._constraints(c, v, n)
        P = P0.copy()
        self._regularise(P, nc)
        res = solve_qp(P, q0, A, b, G, h)
        out = self._finish(c, res, P0, q0, A, b, G, h, cont, P, q0)
        x_tilde = np.hstack([c.rpy - trunk["rpy_body"], c.p_body - trunk["p_body"],
                             (c.p_feet[sw] - c.p_feet_nom[sw]).ravel()])
        self.err = float(x_tilde @ x_tilde)
        self.res = out.primal_res
        out.metrics = self.metrics()
        return out


def _task_stack(c, trunk, sw):
    """Task Jacobian and task-space state/error vectors (clf_controller.py:137-162,
    pc_controller.py:149-183)."""
    t = StepResult()
    if sw:
        t.J = np.vstack([c.J_body] + [c.J_feet[i] for i in sw])
        t.Jdv = np.hstack([c.Jdv_body] + [c.Jdv_feet[i] for i in sw])
    else:
        t.J, t.Jdv = c.J_body, c.Jdv_body
    x = np.hstack([c.rpy, c.p_body, c.p_feet[sw].ravel()])
    xd = np.hstack([c.N @ c.rpyd, c.pd_body, c.pd_feet[sw].ravel()])
    x_nom = np.hstack([trunk["rpy_body"], trunk["p_body"], c.p_feet_nom[sw].ravel()])
    xd_nom = np.hstack([c.N @ trunk["rpyd_body"], trunk["pd_body"], c.pd_feet_nom[sw].ravel()])
    t.xdd_nom = np.hstack([c.N @ trunk["rpydd_body"], trunk["pdd_body"], c.pdd_feet_nom[sw].ravel()])
    t.x_tilde, t.xd_tilde = x - x_nom, xd - xd_nom
    return t


class CLFController(OracleController):
    """clf_controller.py:48-234."""

    def control_law(self, q, v, trunk):
        p = self.p
        q, v = np.asarray(q, float), np.asarray(v, float)
        c = self._common(q, v, trunk)
        sw = [i for i in range(4) if c.swing[i]]
        t = _task_stack(c, trunk, sw)
        eta = np.hstack([t.x_tilde, t.xd_tilde])
        m, nf = len(t.x_tilde), 3 * len(sw)
        Qp = np.diag(np.hstack([p["clf_q_body_rpy"] * np.ones(3), p["clf_q_body_p"] * np.ones(3), p["clf_q_foot_p"] * np.ones(nf)]))
        Qd = np.diag(np.hstack([p["clf_q_body_rpyd"] * np.ones(3), p["clf_q_body_pd"] * np.ones(3), p["clf_q_foot_pd"] * np.ones(nf)]))
        Q = np.block([[Qp, np.zeros((m, m))], [np.zeros((m, m)), Qd]])
        Rm = p["clf_r"] * np.eye(m)
        F = np.block([[np.zeros((m, m)), np.eye(m)], [np.zeros((m, m)), np.zeros((m, m))]])
        Gm = np.vstack([np.zeros((m, m)), np.eye(m)])
        Pl = solve_continuous_are(F, Gm, Q, Rm)          # ContinuousAlgebraicRiccatiEquation :187
        gamma = np.min(np.linalg.eigvals(Q).real) / np.max(np.linalg.eigvals(Pl).real)
        nc = sum(c.contact)
        n = 31 + 3 * nc
        idel = n - 1
        P0, q0 = np.zeros((n, n)), np.zeros(n)
        xdd_des = t.xdd_nom - np.linalg.inv(Rm) @ Gm.T @ Pl @ eta
        P0[:18, :18] += t.J.T @ t.J                      # AddJacobianTypeCost weight 1 :200
        q0[:18] += t.J.T @ (t.Jdv - xdd_des)
        a = 2 * eta @ Pl @ Gm @ t.J                      # AddVdotCost :203
        q0[:18] += a
        P0[idel, idel] += 2 * p["clf_w_delta"]           # AddCost(w*delta'delta) -> Q = 2w
        A, b, G, h, cont = self._constraints(c, v, n)
        V = eta @ Pl @ eta
        row = np.zeros(n)
        row[:18], row[idel] = a, -1.0
        ub = -gamma * V - 2 * eta @ Pl @ F @ eta - 2 * eta @ Pl @ Gm @ (t.Jdv - t.xdd_nom)
        G, h = np.vstack([row[None], G]), np.hstack([ub, h])
        P = P0.copy()
        self._regularise(P, nc)
        res = solve_qp(P, q0, A, b, G, h)
        out = self._finish(c, res, P0, q0, A, b, G, h, cont, P, q0)
        out.delta = res.x[idel]
        self.V = float(V)
        self.err = float(t.x_tilde @ t.x_tilde)
        self.Vdot = float(2 * eta @ Pl @ F @ eta + 2 * eta @ Pl @ Gm @ (t.J @ out.vd + t.Jdv - t.xdd_nom))
        out.metrics = self.metrics()
        out.P_lyap, out.gamma = Pl, gamma
        return out


class PCController(OracleController):
    """pc_controller.py:43-255 with mptc_controller.py:30-57 (AddTaskForceCost).
    `passivity_constraint=False` gives MPTCController.ControlLaw (mptc_controller.py:125-310): same task-force cost and
    gains, no delta, no Vdot rows."""
    passivity_constraint = True

    def control_law(self, q, v, trunk):
        p = self.p
        pl = self.plant
        q, v = np.asarray(q, float), np.asarray(v, float)
        c = self._common(q, v, trunk)
        Cm = pl.coriolis_matrix(q, v)
        sw = [i for i in range(4) if c.swing[i]]
        t = _task_stack(c, trunk, sw)
        Jd = np.vstack([np.zeros((6, 18))] + [pl.frame_jacobian_dot(q, v, pl.foot_frames[i]) for i in sw])
        J = t.J
        Minv = np.linalg.inv(c.M)
        Lam = np.linalg.inv(J @ Minv @ J.T)
        Jbar = Minv @ J.T @ Lam
        Qm = J @ Minv @ Cm - Jd
        nf = 3 * len(sw)
        Kp = np.diag(np.hstack([p["pc_kp_body_rpy"] * np.ones(3), p["pc_kp_body_p"] * np.ones(3), p["pc_kp_foot"] * np.ones(nf)]))
        Kd = np.diag(np.hstack([p["pc_kd_body_rpy"] * np.ones(3), p["pc_kd_body_p"] * np.ones(3), p["pc_kd_foot"] * np.ones(nf)]))
        W = np.diag(np.hstack([p["pc_w_body"] * np.ones(6), p["pc_w_foot"] * np.ones(nf)]))
        f_des = Lam @ t.xdd_nom + Lam @ Qm @ (v - Jbar @ t.xd_tilde) + Jbar.T @ c.tau_g - Kp @ t.x_tilde - Kd @ t.xd_tilde
        nc = sum(c.contact)
        n = 31 + 3 * nc
        idel = n - 1
        A, b, G, h, cont = self._constraints(c, v, n)
        Jc = np.vstack([c.J_feet[i] for i in cont]) if cont else np.zeros((0, 18))
        U = np.hstack([c.S.T, Jc.T])
        P0, q0 = np.zeros((n, n)), np.zeros(n)
        P0[18:18 + 12 + 3 * nc, 18:18 + 12 + 3 * nc] = U.T @ Jbar @ W @ Jbar.T @ U      # AddTaskForceCost
        q0[18:18 + 12 + 3 * nc] = -f_des @ W @ Jbar.T @ U
        row = np.zeros(n)                                 # AddVdotConstraint :14-40
        row[18:18 + 12 + 3 * nc] = t.xd_tilde @ Jbar.T @ U
        row[idel] = -1.0
        ub = t.xd_tilde @ (Jbar.T @ c.tau_g - Lam @ Qm @ (Jbar @ t.xd_tilde - v) + Lam @ t.xdd_nom - Kp @ t.x_tilde)
        row2 = np.zeros(n)                                # delta <= 0 :234-237
        row2[idel] = 1.0
        if self.passivity_constraint:
            G, h = np.vstack([row[None], row2[None], G]), np.hstack([ub, 0.0, h])
        P = P0.copy()
        self._regularise(P, nc, idel, p["reg_f"])
        res = solve_qp(P, q0, A, b, G, h)
        out = self._finish(c, res, P0, q0, A, b, G, h, cont, P, q0)
        out.delta = res.x[idel]
        self.V = float(0.5 * t.xd_tilde @ Lam @ t.xd_tilde + 0.5 * t.x_tilde @ Kp @ t.x_tilde)
        self.err = float(t.x_tilde @ t.x_tilde)
        u = c.S.T @ out.tau + (Jc.T @ np.hstack([out.f[i] for i in cont]) if cont else 0.0)
        fz = Jbar.T @ u
        self.Vdot = float(t.xd_tilde @ (fz - Jbar.T @ c.tau_g + Lam @ Qm @ (Jbar @ t.xd_tilde - v) - Lam @ t.xdd_nom + Kp @ t.x_tilde))
        out.metrics = self.metrics()
        return out


class MPTCController(PCController):
    """mptc_controller.py:125-310 (the unused slack column only carries the tie-break and stays 0)."""
    passivity_constraint = False


class BasicController:
    """BasicController.ControlLaw (basic_controller.py:322-352): joint-space PD about the standing posture, clipped to
    +-150. u = S tau, so only the joint rows matter (MapQDotToVelocity is the identity on them)."""

    def __init__(self, robot="mini_cheetah", dof_order="depth_first", q_nom=None, kp=30.0, kd=1.5, clip=150.0):
        self.plant = robot if isinstance(robot, Plant) else Plant(robot, dof_order)
        self.kp, self.kd, self.clip = kp, kd, clip
        self.q_nom = np.asarray(q_nom, float) if q_nom is not None else np.array(
            [1.0, 0, 0, 0, 0, 0, 0.3] + [0.0, -0.8, 1.6] * 4)

    def control_law(self, q, v, trunk=None):
        S = self.plant.actuation_matrix().T
        tau = np.zeros(18)
        tau[6:] = -self.kp * (np.asarray(q, float)[7:] - self.q_nom[7:]) - self.kd * np.asarray(v, float)[6:]
        return np.clip(S @ tau, -self.clip, self.clip)
