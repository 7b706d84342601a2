"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the Drake MultibodyPlant queries the reference controllers make
(reference controllers/basic_controller.py:101-269). pydrake is an un-vendored,
unpinned dependency of the reference (README.md:9; mid/late-2021 Drake, see
SURVEY.md 8c) and is not installable here, and the reference holds no tests or
golden vectors for this path: **parity unpinned**. What pins this file instead are
the Drake-free known-answer tests of SURVEY.md Appendix D (tests/test_oracle_*.py).

Formulation (deliberately different from the CUDA kernels, which use composite
spatial inertias about the base origin): every link of the *unmerged* URDF tree
(fixed joints and massless frames included) gets world-frame kinematics; inverse
dynamics is Newton-Euler per link projected with the link Jacobians (virtual work);
the mass matrix is built column by column from nv inverse-dynamics passes, which is
what `CalcMassMatrixViaInverseDynamics` (basic_controller.py:110) does.

Drake conventions restated:
  q = [qw qx qy qz | p_W | theta],  v = [omega_W | pdot_W | thetadot]   (A.1)
  CalcBiasTerm = ID(q, v, vdot=0) without gravity; the controller's tau_g is
  -CalcGravityGeneralizedForces (basic_controller.py:112)                 (A.3)
  Jacobians wrt v, expressed in world; bias = Jdot*v = classical point
  acceleration at vdot = 0 (basic_controller.py:180-195)                  (A.6)
"""
from __future__ import annotations

import json
import math
from pathlib import Path

import numpy as np

ROBOT_DIR = Path(__file__).resolve().parents[1] / "quadruped_drake_b200" / "robots"
GRAVITY = np.array([0.0, 0.0, -9.81])


def skew(a):
    return np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])


def rpy_matrix(rpy):
    """R = Rz(y) Ry(p) Rx(r) (Drake RollPitchYaw)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def rpy_from_matrix(R):
    """RollPitchYaw(RotationMatrix).vector() away from gimbal lock."""
    pitch = math.atan2(-R[2, 0], math.hypot(R[0, 0], R[1, 0]))
    return np.array([math.atan2(R[2, 1], R[2, 2]), pitch, math.atan2(R[1, 0], R[0, 0])])


def rpy_rate_matrix(rpy):
    """N with omega_W = N(rpy) rpydot (CalcAngularVelocityInParentFromRpyDt)."""
    _, p, y = rpy
    cp, sp, cy, sy = math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, -sy, 0.0], [sy * cp, cy, 0.0], [-sp, 0.0, 1.0]])


def quat_to_matrix(qw, qx, qy, qz):
    n = math.sqrt(qw * qw + qx * qx + qy * qy + qz * qz)
    w, x, y, z = qw / n, qx / n, qy / n, qz / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def axis_angle_matrix(a, th):
    a = np.asarray(a, float)
    a = a / np.linalg.norm(a)
    K = skew(a)
    return np.eye(3) + math.sin(th) * K + (1.0 - math.cos(th)) * (K @ K)


class Plant:
    """Floating-base tree read straight from the neutral description (no welding)."""

    def __init__(self, robot="mini_cheetah", dof_order="depth_first"):
        desc = json.loads((ROBOT_DIR / f"{robot}.json").read_text()) if isinstance(robot, str) else robot
        self.name = desc["name"]
        links = {l["name"]: l for l in desc["links"]}
        pj = {j["child"]: j for j in desc["joints"]}
        base = desc["base_link"]
        # topological order = order of appearance, parents first
        order, todo = [base], [l["name"] for l in desc["links"] if l["name"] != base]
        while todo:
            progressed = False
            for n in list(todo):
                if pj[n]["parent"] in order:
                    order.append(n)
                    todo.remove(n)
                    progressed = True
            if not progressed:
                raise ValueError("disconnected links: %s" % todo)
        self.names = order
        self.index = {n: i for i, n in enumerate(order)}
        self.parent = [-1] + [self.index[pj[n]["parent"]] for n in order[1:]]
        self.joint = [None] + [pj[n] for n in order[1:]]
        self.mass = np.array([links[n]["mass"] for n in order])
        self.com = np.array([links[n]["com"] for n in order])
        self.Icom = []
        for n in order:
            xx, yy, zz, xy, xz, yz = links[n]["inertia_com"]
            self.Icom.append(np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]]))
        self.depth = [0] * len(order)
        for i in range(1, len(order)):
            self.depth[i] = self.depth[self.parent[i]] + 1
        rev = [i for i in range(1, len(order)) if self.joint[i]["type"] == "revolute"]  # file order of links
        # Drake numbers dofs in joint order; both reference URDFs declare joints leg by leg.
        jorder = [j["name"] for j in desc["joints"] if j["type"] == "revolute"]
        by_name = {self.joint[i]["name"]: i for i in rev}
        if dof_order == "depth_first":
            seq = [by_name[n] for n in jorder]
        elif dof_order == "breadth_first":
            seq = sorted((by_name[n] for n in jorder), key=lambda i: (self._moving_depth(i), jorder.index(self.joint[i]["name"])))
        else:
            seq = [by_name[n] for n in dof_order]
        self.vidx = [None] * len(order)
        for pos, i in enumerate(seq):
            self.vidx[i] = 6 + pos
        self.nv, self.nq = 6 + len(seq), 7 + len(seq)
        self.actuated = [by_name[n] for n in desc["actuated_joints"]]
        self.effort = np.array([self.joint[i]["effort"] for i in self.actuated])
        self.base_link = base
        self.foot_frames = list(desc["foot_frames"])

    def _moving_depth(self, i):
        d = 0
        while i > 0:
            if self.joint[i]["type"] == "revolute":
                d += 1
            i = self.parent[i]
        return d

    # ------------------------------------------------------------------ kinematics
    def kinematics(self, q, v, vd=None):
        """World pose, angular velocity/acceleration and origin velocity/classical
        acceleration of every link; `vd` = generalized acceleration in Drake coordinates."""
        q, v = np.asarray(q, float), np.asarray(v, float)
        vd = np.zeros(self.nv) if vd is None else np.asarray(vd, float)
        n = len(self.names)
        R, p, w, al, vo, ao, ax = [None] * n, [None] * n, [None] * n, [None] * n, [None] * n, [None] * n, [None] * n
        R[0] = quat_to_matrix(*q[0:4])
        p[0] = q[4:7].copy()
        w[0], vo[0], al[0], ao[0] = v[0:3].copy(), v[3:6].copy(), vd[0:3].copy(), vd[3:6].copy()
        for i in range(1, n):
            j, pa = self.joint[i], self.parent[i]
            r = R[pa] @ np.array(j["xyz"])
            RJ = R[pa] @ rpy_matrix(j["rpy"])
            p[i] = p[pa] + r
            vo[i] = vo[pa] + np.cross(w[pa], r)
            ao[i] = ao[pa] + np.cross(al[pa], r) + np.cross(w[pa], np.cross(w[pa], r))
            if j["type"] == "revolute":
                k = self.vidx[i]
                a_w = RJ @ (np.array(j["axis"]) / np.linalg.norm(j["axis"]))
                R[i] = RJ @ axis_angle_matrix(j["axis"], q[k + 1])
                w[i] = w[pa] + a_w * v[k]
                al[i] = al[pa] + a_w * vd[k] + np.cross(w[pa], a_w) * v[k]
                ax[i] = a_w
            else:
                R[i], w[i], al[i] = RJ, w[pa].copy(), al[pa].copy()
        return dict(R=R, p=p, w=w, al=al, vo=vo, ao=ao, axis=ax)

    def _jac_w(self, kin, i):
        J = np.zeros((3, self.nv))
        J[:, 0:3] = np.eye(3)
        while i > 0:
            if self.vidx[i] is not None:
                J[:, self.vidx[i]] = kin["axis"][i]
            i = self.parent[i]
        return J

    def _jac_v(self, kin, i, pt):
        J = np.zeros((3, self.nv))
        J[:, 0:3] = -skew(pt - kin["p"][0])
        J[:, 3:6] = np.eye(3)
        while i > 0:
            if self.vidx[i] is not None:
                J[:, self.vidx[i]] = np.cross(kin["axis"][i], pt - kin["p"][i])
            i = self.parent[i]
        return J

    # -------------------------------------------------------------- inverse dynamics
    def inverse_dynamics(self, q, v, vd, gravity=True):
        """tau with  M vd + C v + tau_g(controller sign) = tau."""
        kin = self.kinematics(q, v, vd)
        tau = np.zeros(self.nv)
        g = GRAVITY if gravity else np.zeros(3)
        for i in range(len(self.names)):
            m = self.mass[i]
            if m == 0.0:
                continue
            rc = kin["R"][i] @ self.com[i]
            w, al = kin["w"][i], kin["al"][i]
            a_c = kin["ao"][i] + np.cross(al, rc) + np.cross(w, np.cross(w, rc))
            Iw = kin["R"][i] @ self.Icom[i] @ kin["R"][i].T
            F = m * (a_c - g)
            N = Iw @ al + np.cross(w, Iw @ w)
            tau += self._jac_v(kin, i, kin["p"][i] + rc).T @ F + self._jac_w(kin, i).T @ N
        return tau

    def mass_matrix(self, q):
        """CalcMassMatrixViaInverseDynamics: column j = ID(q, 0, e_j) without gravity."""
        z = np.zeros(self.nv)
        M = np.zeros((self.nv, self.nv))
        for j in range(self.nv):
            e = np.zeros(self.nv)
            e[j] = 1.0
            M[:, j] = self.inverse_dynamics(q, z, e, gravity=False)
        return M

    def bias_term(self, q, v):
        """CalcBiasTerm: C(q,v) v."""
        return self.inverse_dynamics(q, v, np.zeros(self.nv), gravity=False)

    def gravity_term(self, q):
        """Controller-sign tau_g = -CalcGravityGeneralizedForces (basic_controller.py:112)."""
        z = np.zeros(self.nv)
        return self.inverse_dynamics(q, z, z, gravity=True)

    def actuation_matrix(self):
        """MakeActuationMatrix: nv x nu, one 1 per column at the joint's velocity index."""
        B = np.zeros((self.nv, len(self.actuated)))
        for a, i in enumerate(self.actuated):
            B[self.vidx[i], a] = 1.0
        return B

    def calc_dynamics(self, q, v):
        """BasicController.CalcDynamics (basic_controller.py:101-115): M, Cv, tau_g, S."""
        return self.mass_matrix(q), self.bias_term(q, v), self.gravity_term(q), self.actuation_matrix().T

    def coriolis_matrix(self, q, v):
        """CalcCoriolisMatrix (basic_controller.py:117-132): 0.5 d(Cv)/dv. Cv is a homogeneous
        quadratic form in v, so the derivative is exact by polarisation."""
        v = np.asarray(v, float)
        b0 = self.bias_term(q, v)
        Cm = np.zeros((self.nv, self.nv))
        for j in range(self.nv):
            e = np.zeros(self.nv)
            e[j] = 1.0
            Cm[:, j] = 0.5 * (self.bias_term(q, v + e) - b0 - self.bias_term(q, e))
        return Cm

    # ------------------------------------------------------------------ frame queries
    def frame_position_quantities(self, q, v, frame):
        """CalcFramePositionQuantities (basic_controller.py:173-196): p, J (3 x nv), Jdot*v."""
        kin = self.kinematics(q, v)
        i = self.index[frame]
        pt = kin["p"][i]
        return pt.copy(), self._jac_v(kin, i, pt), kin["ao"][i].copy()

    def frame_pose_quantities(self, q, v, frame):
        """CalcFramePoseQuantities (basic_controller.py:246-269): (R, p), J (6 x nv, rows
        [angular; linear]), bias spatial acceleration (6)."""
        kin = self.kinematics(q, v)
        i = self.index[frame]
        J = np.vstack([self._jac_w(kin, i), self._jac_v(kin, i, kin["p"][i])])
        return (kin["R"][i].copy(), kin["p"][i].copy()), J, np.hstack([kin["al"][i], kin["ao"][i]])

    def frame_jacobian_dot(self, q, v, frame):
        """CalcFrameJacobianDot (basic_controller.py:198-220): d/dt of the translational
        Jacobian along qdot = N(q) v, column by column."""
        kin = self.kinematics(q, v)
        i = self.index[frame]
        pt, vpt = kin["p"][i], kin["vo"][i]
        Jd = np.zeros((3, self.nv))
        Jd[:, 0:3] = -skew(vpt - kin["vo"][0])
        k = i
        while k > 0:
            if self.vidx[k] is not None:
                a, pa = kin["axis"][k], self.parent[k]
                adot = np.cross(kin["w"][pa], a)
                Jd[:, self.vidx[k]] = np.cross(adot, pt - kin["p"][k]) + np.cross(a, vpt - kin["vo"][k])
            k = self.parent[k]
        return Jd

    def total_mass(self):
        return float(self.mass.sum())
