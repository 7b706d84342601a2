"""KKT certificate of the *reference* QP for a whole batch (test infrastructure, vectorised numpy).

Given the step outputs x = [vd; tau; f] (+ delta for CLF) and the multipliers `lam` the library exports, this rebuilds
the cost and constraints of the reference ControlLaw (reference
controllers/inverse_dynamics_controller.py:25-101,116-221, clf_controller.py:137-221) for every instance of the batch
from the dynamics terms (M, Cv, tau_g, J, Jdot v, p: `wbc_dynamics`, itself parity-tested against the oracle) and checks

    stationarity   P x + c + G' lam in range(A')         (equality multipliers are eliminated by projection)
    dual           lam >= 0
    complementary  lam_i (h_i - G_i x) = 0
    primal         A x = b,  G x <= h

so that a feasible but sub-optimal torque fails. The declared tie-break (reg_f / reg_tau) is part of P.
Size independent: used at the full BASELINE sizes where the per-instance Python oracle would take hours.
"""
from __future__ import annotations

import numpy as np

DEFAULTS = dict(id_kp_body_p=500.0, id_kd_body_p=50.0, id_kp_body_rpy=500.0, id_kd_body_rpy=50.0, id_kp_foot=100.0,
                id_kd_foot=20.0, id_w_body=10.0, id_w_foot=1.0,
                clf_q_body_p=5000.0, clf_q_body_pd=200.0, clf_q_body_rpy=5000.0, clf_q_body_rpyd=200.0,
                clf_q_foot_p=200.0, clf_q_foot_pd=20.0, clf_r=1.0, clf_w_delta=1000.0,
                mu=0.7, contact_damping=100.0, reg_f=1e-6, reg_tau=0.0, torque_limits=0)


def quat_to_rpy(quat):
    w, x, y, z = (quat[:, i] / np.linalg.norm(quat, axis=1) for i in range(4))
    r00, r10, r20 = 1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)
    r21, r22 = 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)
    return np.stack([np.arctan2(r21, r22), np.arctan2(-r20, np.hypot(r00, r10)), np.arctan2(r10, r00)], axis=1)


def rate_matrix(rpy):
    """N with omega_W = N(rpy) rpydot."""
    n = len(rpy)
    cp, sp, cy, sy = np.cos(rpy[:, 1]), np.sin(rpy[:, 1]), np.cos(rpy[:, 2]), np.sin(rpy[:, 2])
    N = np.zeros((n, 3, 3))
    N[:, 0, 0], N[:, 0, 1] = cy * cp, -sy
    N[:, 1, 0], N[:, 1, 1] = sy * cp, cy
    N[:, 2, 0], N[:, 2, 2] = -sp, 1.0
    return N


def care_channel(qp, qd, r):
    """Scalar double-integrator CARE: p11, p12, p22."""
    p12 = np.sqrt(qp * r)
    p22 = np.sqrt(r * (qd + 2 * p12))
    return p12 * p22 / r, p12, p22


def _task_quantities(d, q, v, traj):
    n = len(q)
    rpy = quat_to_rpy(q[:, 0:4])
    N = rate_matrix(rpy)
    rpyd = np.linalg.solve(N, v[:, 0:3, None])[:, :, 0]
    Jb = np.zeros((n, 6, 18))
    Jb[:, np.arange(6), np.arange(6)] = 1.0
    return rpy, N, rpyd, Jb


def reference_qp(kind, d, model, q, v, traj, contact, params=None):
    """Cost and constraints of the reference QP over x = [vd(18); tau(12); f(12: LF RF LH RH); delta] for every instance.
    Swing feet keep their three force variables, pinned by the equality f_k = 0 (their multipliers are free, so the
    Lagrangian conditions of the remaining variables are those of the reference program without them)."""
    p = dict(DEFAULTS)
    p.update(params or {})
    n = len(q)
    stance = contact.astype(bool)
    sw = (~stance).astype(float)
    rpy, N, rpyd, Jb = _task_quantities(d, q, v, traj)
    J, Jdv, pf = d["J_feet"], d["Jdv_feet"], d["p_feet"]
    pdf = np.einsum("nkij,nj->nki", J, v)
    nx = 43
    P, c = np.zeros((n, nx, nx)), np.zeros((n, nx))
    G_rows, h_rows, slots = [], [], []
    tr = traj
    p_nom_f, pd_nom_f, pdd_nom_f = tr[:, 18:30].reshape(n, 4, 3), tr[:, 30:42].reshape(n, 4, 3), tr[:, 42:54].reshape(n, 4, 3)
    if kind == "id":
        rdd = tr[:, 15:18] - p["id_kp_body_rpy"] * (rpy - tr[:, 9:12]) - p["id_kd_body_rpy"] * (rpyd - tr[:, 12:15])
        add = tr[:, 6:9] - p["id_kp_body_p"] * (q[:, 4:7] - tr[:, 0:3]) - p["id_kd_body_p"] * (v[:, 3:6] - tr[:, 3:6])
        ab = np.concatenate([np.einsum("nij,nj->ni", N, rdd), add], axis=1)
        P[:, :18, :18] += p["id_w_body"] * np.einsum("nri,nrj->nij", Jb, Jb)
        c[:, :18] += p["id_w_body"] * np.einsum("nri,nr->ni", Jb, -ab)
        a_s = pdd_nom_f - p["id_kp_foot"] * (pf - p_nom_f) - p["id_kd_foot"] * (pdf - pd_nom_f)
        P[:, :18, :18] += p["id_w_foot"] * np.einsum("nk,nkri,nkrj->nij", sw, J, J)
        c[:, :18] += p["id_w_foot"] * np.einsum("nk,nkri,nkr->ni", sw, J, Jdv - a_s)
    elif kind == "clf":
        ch = [care_channel(p["clf_q_body_rpy"], p["clf_q_body_rpyd"], p["clf_r"]),
              care_channel(p["clf_q_body_p"], p["clf_q_body_pd"], p["clf_r"]),
              care_channel(p["clf_q_foot_p"], p["clf_q_foot_pd"], p["clf_r"])]
        # task rows: 0-2 rpy, 3-5 position, 6-17 feet (masked by swing)
        xt = np.concatenate([rpy - tr[:, 9:12], q[:, 4:7] - tr[:, 0:3], (pf - p_nom_f).reshape(n, 12)], axis=1)
        xdt = np.concatenate([v[:, 0:3] - np.einsum("nij,nj->ni", N, tr[:, 12:15]), v[:, 3:6] - tr[:, 3:6],
                              (pdf - pd_nom_f).reshape(n, 12)], axis=1)
        xddn = np.concatenate([np.einsum("nij,nj->ni", N, tr[:, 15:18]), tr[:, 6:9], pdd_nom_f.reshape(n, 12)], axis=1)
        Jt = np.concatenate([Jb, J.reshape(n, 12, 18)], axis=1)
        Jdvt = np.concatenate([np.zeros((n, 6)), Jdv.reshape(n, 12)], axis=1)
        mask = np.concatenate([np.ones((n, 6)), np.repeat(sw, 3, axis=1)], axis=1)
        p11 = np.array([ch[0][0]] * 3 + [ch[1][0]] * 3 + [ch[2][0]] * 12)
        p12 = np.array([ch[0][1]] * 3 + [ch[1][1]] * 3 + [ch[2][1]] * 12)
        p22 = np.array([ch[0][2]] * 3 + [ch[1][2]] * 3 + [ch[2][2]] * 12)
        kap = mask * (p12 * xt + p22 * xdt)                                 # G' P eta
        xdd_des = xddn - kap / p["clf_r"]
        a = 2.0 * np.einsum("nr,nri->ni", kap, Jt)
        P[:, :18, :18] += np.einsum("nr,nri,nrj->nij", mask, Jt, Jt)
        c[:, :18] += np.einsum("nr,nri,nr->ni", mask, Jt, Jdvt - xdd_des) + a
        P[:, 42, 42] += 2.0 * p["clf_w_delta"]
        V = (mask * (p11 * xt * xt + 2 * p12 * xt * xdt + p22 * xdt * xdt)).sum(axis=1)
        PF = (mask * (p11 * xt + p12 * xdt) * xdt).sum(axis=1)
        lmax = [0.5 * (a11 + a22) + np.sqrt(0.25 * (a11 - a22) ** 2 + a12 * a12) for a11, a12, a22 in ch]
        qmin_body = min(p["clf_q_body_rpy"], p["clf_q_body_rpyd"], p["clf_q_body_p"], p["clf_q_body_pd"])
        qmin_all = min(qmin_body, p["clf_q_foot_p"], p["clf_q_foot_pd"])
        any_sw = sw.sum(axis=1) > 0
        gamma = np.where(any_sw, qmin_all / max(lmax), qmin_body / max(lmax[0], lmax[1]))
        ub = -gamma * V - 2.0 * PF - 2.0 * (kap * (Jdvt - xddn)).sum(axis=1)
        row = np.zeros((n, nx))
        row[:, :18], row[:, 42] = a, -1.0
        G_rows.append(row[:, None, :]); h_rows.append(ub[:, None]); slots.append([16])
    else:
        raise ValueError(kind)
    # tie-break
    P[:, np.arange(18, 30), np.arange(18, 30)] += p["reg_tau"]
    P[:, np.arange(30, 42), np.arange(30, 42)] += p["reg_f"]
    if kind != "clf":
        P[:, 42, 42] += 1.0                       # unused slack column: pinned at 0 by its own cost
    # equalities: dynamics (18), per foot either the no-slip rows (stance) or f_k = 0 (swing)
    B = model.actuation_matrix()
    A, b = np.zeros((n, 30, nx)), np.zeros((n, 30))
    A[:, :18, :18] = d["M"]
    A[:, :18, 18:30] = -B
    for k in range(4):
        st = stance[:, k].astype(float)[:, None, None]
        A[:, :18, 30 + 3 * k:33 + 3 * k] = -st * np.transpose(J[:, k], (0, 2, 1))
        A[:, 18 + 3 * k:21 + 3 * k, :18] = st * J[:, k]
        A[:, 18 + 3 * k + np.arange(3), 30 + 3 * k + np.arange(3)] = 1.0 - stance[:, k].astype(float)[:, None]
        b[:, 18 + 3 * k:21 + 3 * k] = stance[:, k, None] * (-Jdv[:, k] - p["contact_damping"] * pdf[:, k])
    b[:, :18] = -d["Cv"] - d["tau_g"]
    # inequalities in the lam layout of wbc.h
    mu = p["mu"]
    A_i = np.array([[1, 0, -mu], [-1, 0, -mu], [0, 1, -mu], [0, -1, -mu]], float)
    fr = np.zeros((n, 16, nx))
    for k in range(4):
        fr[:, 4 * k:4 * k + 4, 30 + 3 * k:33 + 3 * k] = A_i[None] * stance[:, k, None, None]
    G_rows.insert(0, fr); h_rows.insert(0, np.zeros((n, 16))); slots.insert(0, list(range(16)))
    if p["torque_limits"]:
        eff = model.effort[np.argsort(model.act_index)]          # actuator order
        tl = np.zeros((n, 24, nx))
        tl[:, np.arange(12), 18 + np.arange(12)] = 1.0
        tl[:, 12 + np.arange(12), 18 + np.arange(12)] = -1.0
        G_rows.append(tl); h_rows.append(np.tile(np.concatenate([eff, eff]), (n, 1))); slots.append(list(range(18, 42)))
    return P, c, A, b, np.concatenate(G_rows, axis=1), np.concatenate(h_rows, axis=1), sum(slots, [])


def certificate(kind, d, model, q, v, traj, contact, out, params=None):
    """-> dict of per-instance residuals: stationarity (relative), dual (min lam), comp, eq, ineq."""
    P, c, A, b, G, h, slots = reference_qp(kind, d, model, q, v, traj, contact, params)
    n = len(q)
    x = np.zeros((n, 43))
    x[:, :18], x[:, 18:30], x[:, 30:42] = out.vd, out.tau, out.f.reshape(n, 12)
    if kind == "clf":
        x[:, 42] = out.qp_info[:, 2]
    lam = out.lam[:, slots]
    grad = np.einsum("nij,nj->ni", P, x) + c + np.einsum("nri,nr->ni", G, lam)
    Q, _ = np.linalg.qr(np.transpose(A, (0, 2, 1)))                       # range(A') per instance
    resid = grad - np.einsum("nir,nr->ni", Q, np.einsum("nir,ni->nr", Q, grad))
    gscale = np.maximum(1.0, np.abs(np.einsum("nij,nj->ni", P, x) + c).max(axis=1))
    slack = h - np.einsum("nri,ni->nr", G, x)
    return {"stationarity": np.abs(resid).max(axis=1) / gscale,
            "dual": lam.min(axis=1),
            "comp": np.abs(lam * slack).max(axis=1) / np.maximum(1.0, np.abs(lam).max(axis=1)),
            "eq": np.abs(np.einsum("nri,ni->nr", A, x) - b).max(axis=1),
            "ineq": np.maximum(0.0, -slack).max(axis=1),
            "n_active": (lam > 0).sum(axis=1)}
