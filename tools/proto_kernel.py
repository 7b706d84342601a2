#!/usr/bin/env python
"""Numpy sketch of the FIRST step kernel's algorithm (one instance at a time), kept as a second derivation of the dynamics.

Not shipped and not an oracle. Its dynamics part (composite spatial inertias about the base origin, world axes) is the
formulation `dynamics_phase` in csrc/wbc_device.cuh still implements, and tests/test_oracle_dynamics.py ("two
derivations, one answer") checks it against oracle/dynamics.py, which uses projected Newton-Euler on the unmerged tree. The
QP part below describes the round-1 textbook Goldfarb-Idnani iteration on J; the kernel has since moved to the W = Y J form
with explicit R^-1 rows - the host warp emulator (tests/emu) is what runs the current device code on a CPU.

  1. dynamics: composite spatial inertias about the base origin, world axes
  2. equality elimination: Gauss-Jordan with column pivoting on [A | b], z = z0 + Z w
  3. reduced cost/inequalities, 4. Goldfarb-Idnani dual active set on w, 5. tau recovery
"""
import numpy as np


def skew(a):
    return np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])


def rodrigues(a, th):
    K = skew(a)
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def quat_R(qw, qx, qy, qz):
    n = np.sqrt(qw * qw + qx * qx + qy * qy + qz * qz)
    w, x, y, z = qw / n, qx / n, qy / n, qz / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def crm(v):   # spatial motion cross product, [angular; linear]
    w, u = v[:3], v[3:]
    return np.block([[skew(w), np.zeros((3, 3))], [skew(u), skew(w)]])


def dynamics(model, q, v, gravity_in_bias=True):
    """Internal joint order k = 3*leg + j. Returns dict with M (18x18 internal order), h = Cv (+ tau_g),
    taug, foot p / rho / L (3x3 leg Jacobian) / Jdv / velocity, R, everything about P = base origin."""
    R0 = quat_R(*q[0:4])
    P = q[4:7]
    th = np.array([q[1 + model.v_index[k]] for k in range(12)])
    thd = np.array([v[model.v_index[k]] for k in range(12)])
    g = model.gravity
    wb, vb = v[0:3], v[3:6]

    def spatial_inertia(m, c, Ic):
        """6x6 about P for a body with mass m, CoM offset c (from P, world axes), inertia Ic about CoM."""
        S = skew(c)
        return np.block([[Ic + m * S @ S.T, m * S], [m * S.T, m * np.eye(3)]])

    Ib = spatial_inertia(model.mass[0], R0 @ model.com[0], R0 @ sym6(model.inertia_com[0]) @ R0.T)
    vbase = np.hstack([wb, vb])
    abase = np.hstack([np.zeros(3), -np.cross(wb, vb) - (g if gravity_in_bias else 0)])
    M = np.zeros((18, 18))
    h = np.zeros(18)
    Itot = Ib.copy()
    # v x* (I v) = -crm(v)^T (I v)
    ftot = Ib @ abase - crm(vbase).T @ (Ib @ vbase)
    out = dict(p=np.zeros((4, 3)), rho=np.zeros((4, 3)), L=np.zeros((4, 3, 3)), Jdv=np.zeros((4, 3)), vf=np.zeros((4, 3)),
               Ld=np.zeros((4, 3, 3)))
    mc = [model.mass[0], model.mass[0] * (R0 @ model.com[0])]
    for leg in range(4):
        R, rho = R0.copy(), np.zeros(3)
        vel, acc = vbase.copy(), abase.copy()
        S, I, f, vl = [], [], [], []
        axes, origins, omegas, vorig = [], [], [], []
        for j in range(3):
            k = 3 * leg + j
            rho = rho + R @ model.joint_xyz[k]
            a = R @ model.joint_axis[k]
            R = R @ rodrigues(model.joint_axis[k], th[k])
            Sj = np.hstack([a, np.cross(rho, a)])
            # velocity of the joint origin (needed for Jdot): parent's spatial velocity evaluated at rho
            vorig.append(vel[3:] + np.cross(vel[:3], rho))
            omegas.append(vel[:3].copy())           # parent angular velocity
            acc = acc + crm(vel) @ Sj * thd[k]
            vel = vel + Sj * thd[k]
            Ij = spatial_inertia(model.mass[k + 1], rho + R @ model.com[k + 1], R @ sym6(model.inertia_com[k + 1]) @ R.T)
            fj = Ij @ acc - crm(vel).T @ (Ij @ vel)
            S.append(Sj); I.append(Ij); f.append(fj); vl.append(vel.copy())
            axes.append(a); origins.append(rho.copy())
        # composites (leaf to root)
        Ic = [None] * 3
        fc = [None] * 3
        Ic[2], fc[2] = I[2], f[2]
        Ic[1], fc[1] = I[1] + Ic[2], f[1] + fc[2]
        Ic[0], fc[0] = I[0] + Ic[1], f[0] + fc[1]
        Itot += Ic[0]
        ftot += fc[0]
        for j in range(3):
            k = 6 + 3 * leg + j
            F = Ic[j] @ S[j]
            M[0:6, k] = F
            M[k, 0:6] = F
            for i in range(j + 1):
                ki = 6 + 3 * leg + i
                M[ki, k] = M[k, ki] = S[i] @ F
            h[k] = S[j] @ fc[j]
        # foot
        rf = rho + R @ model.foot_xyz[leg]
        out["rho"][leg] = rf
        out["p"][leg] = P + rf
        for j in range(3):
            out["L"][leg][:, j] = np.cross(axes[j], rf - origins[j])
        w2, vP2 = vel[:3], vel[3:]
        vfoot = vP2 + np.cross(w2, rf)
        out["vf"][leg] = vfoot
        out["Jdv"][leg] = acc[3:] + np.cross(acc[:3], rf) + np.cross(w2, vfoot) + (g if gravity_in_bias else 0)
        for j in range(3):                         # Jdot columns of the leg joints (PC only)
            adot = np.cross(omegas[j], axes[j])
            out["Ld"][leg][:, j] = np.cross(adot, rf - origins[j]) + np.cross(axes[j], vfoot - vorig[j])
    M[0:6, 0:6] = Itot
    h[0:6] = ftot
    out.update(M=M, h=h, R0=R0, P=P, Itot=Itot)
    return out


def sym6(v6):
    xx, yy, zz, xy, xz, yz = v6
    return np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])


def foot_jacobian(dyn, leg):
    """3 x 18 in internal order."""
    J = np.zeros((3, 18))
    J[:, 0:3] = -skew(dyn["rho"][leg])
    J[:, 3:6] = np.eye(3)
    J[:, 6 + 3 * leg:9 + 3 * leg] = dyn["L"][leg]
    return J


# ------------------------------------------------------------------------------- reduction
def gauss_jordan(A, b, tol=1e-11):
    """Row by row, pivot = largest remaining |entry| of the row (column pivoting).
    Returns pivot column per row (-1 if the row vanished), reduced [A|b], rank-deficiency flag."""
    A = A.copy()
    b = b.copy()
    m, n = A.shape
    used = np.zeros(n, bool)
    pc = -np.ones(m, int)
    flag = 0
    for r in range(m):
        cand = np.where(used, 0.0, np.abs(A[r]))
        c = int(np.argmax(cand))
        scale = np.abs(A[r]).max()
        if cand[c] <= tol * max(1.0, scale):
            flag = 1
            continue
        pc[r] = c
        used[c] = True
        piv = A[r, c]
        A[r] /= piv
        b[r] /= piv
        for i in range(m):
            if i != r:
                fct = A[i, c]
                A[i] -= fct * A[r]
                b[i] -= fct * b[r]
    return pc, A, b, flag


def nullspace(A, b):
    pc, Ar, br, flag = gauss_jordan(A, b)
    n = A.shape[1]
    free = [c for c in range(n) if c not in set(pc[pc >= 0])]
    Z = np.zeros((n, len(free)))
    z0 = np.zeros(n)
    for w, c in enumerate(free):
        Z[c, w] = 1.0
    for r, c in enumerate(pc):
        if c >= 0:
            z0[c] = br[r]
            Z[c, :] = -Ar[r, free]
    return z0, Z, flag


# ------------------------------------------------------------------------- Goldfarb-Idnani
def gi_solve(H, g, G, hv, max_iter=100, status=None):
    """min 1/2 w'Hw + g'w  s.t.  G w <= hv.  Returns w, multipliers u (per constraint), iters, flag."""
    n = H.shape[0]
    mi = G.shape[0]
    L = np.linalg.cholesky(H)
    J = np.linalg.inv(L).T            # J J' = H^-1
    x = -J @ (J.T @ g)
    R = np.zeros((n, n))
    act = []                          # active constraint ids in order
    u = []                            # multipliers of active constraints
    q = 0
    it = 0
    flag = 0
    excluded = np.zeros(mi, bool)
    while True:
        s = hv - G @ x
        tol = 1e-10 * (1.0 + np.abs(hv) + np.abs(G) @ np.abs(x))
        viol = np.where(excluded, 0.0, s + tol)
        for a in act:
            viol[a] = 0.0
        if mi == 0 or viol.min() >= 0.0:
            break
        p = int(np.argmin(np.where(viol < 0, s, np.inf)))
        npv = -G[p]
        up = 0.0
        while True:                   # step 2
            it += 1
            if it > max_iter:
                flag |= 1
                break
            d = J.T @ npv
            z = J[:, q:] @ d[q:]
            r = np.linalg.solve(R[:q, :q], d[:q]) if q else np.zeros(0)
            t1, l = np.inf, -1
            for k in range(q):
                if r[k] > 0 and u[k] / r[k] < t1:
                    t1, l = u[k] / r[k], k
            zn = z @ npv              # = |d[q:]|^2 >= 0
            sp = hv[p] - G[p] @ x
            dscale = d @ d
            if zn > 1e-14 * max(dscale, 1e-300):
                t2 = -sp / zn
            else:
                t2 = np.inf
            t = min(t1, t2)
            if not np.isfinite(t):
                flag |= 2             # infeasible
                break
            if not np.isfinite(t2):   # dual step only
                for k in range(q):
                    u[k] -= t * r[k]
                up += t
                J, R, act, u, q = _drop(J, R, act, u, q, l)
                continue
            x = x + t * z
            for k in range(q):
                u[k] -= t * r[k]
            up += t
            if t == t2:               # full step: add p
                J, R, ok = _add(J, R, d, q)
                act.append(p)
                u.append(up)
                q += 1
                break
            J, R, act, u, q = _drop(J, R, act, u, q, l)
        if flag:
            break
    lam = np.zeros(mi)
    for a, ua in zip(act, u):
        lam[a] = ua
    return x, lam, it, flag


def _add(J, R, d, q):
    n = J.shape[0]
    v = d[q:].copy()
    nv = np.linalg.norm(v)
    R = R.copy()
    J = J.copy()
    if n - q > 1:
        alpha = -np.copysign(nv, v[0]) if v[0] != 0 else -nv
        v[0] -= alpha
        vv = v @ v
        if vv > 0:
            J[:, q:] -= np.outer(J[:, q:] @ v, 2.0 * v / vv)
        R[:q, q] = d[:q]
        R[q, q] = alpha
    else:
        R[:q, q] = d[:q]
        R[q, q] = v[0]
    return J, R, True


def _drop(J, R, act, u, q, l):
    J, R = J.copy(), R.copy()
    n = J.shape[0]
    # remove column l
    R[:, l:q - 1] = R[:, l + 1:q]
    R[:, q - 1] = 0.0
    for k in range(l, q - 1):
        a, b = R[k, k], R[k + 1, k]
        rr = np.hypot(a, b)
        if rr == 0.0:
            continue
        c, s = a / rr, b / rr
        Rk, Rk1 = R[k].copy(), R[k + 1].copy()
        R[k], R[k + 1] = c * Rk + s * Rk1, -s * Rk + c * Rk1
        Jk, Jk1 = J[:, k].copy(), J[:, k + 1].copy()
        J[:, k], J[:, k + 1] = c * Jk + s * Jk1, -s * Jk + c * Jk1
    act = act[:l] + act[l + 1:]
    u = u[:l] + u[l + 1:]
    return J, R, act, u, q - 1


# ----------------------------------------------------------------------------- ID step
def rpy_from_R(R):
    return np.array([np.arctan2(R[2, 1], R[2, 2]), np.arctan2(-R[2, 0], np.hypot(R[0, 0], R[1, 0])),
                     np.arctan2(R[1, 0], R[0, 0])])


def rpy_N(rpy):
    _, p, y = rpy
    cp, sp, cy, sy = np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, -sy, 0.0], [sy * cp, cy, 0.0], [-sp, 0.0, 1.0]])


def build_equalities(model, prm, dyn, v_int, contact):
    """A z = b for z = [a_b(6), a_j(12), f(3 nc)] with tau eliminated: 6 base rows + 3 rows per stance foot."""
    cont = [i for i in range(4) if contact[i]]
    nc = len(cont)
    n = 18 + 3 * nc
    A = np.zeros((6 + 3 * nc, n))
    b = np.zeros(6 + 3 * nc)
    A[0:6, 0:18] = dyn["M"][0:6, :]
    b[0:6] = -dyn["h"][0:6]
    for j, i in enumerate(cont):
        Ji = foot_jacobian(dyn, i)
        A[0:6, 18 + 3 * j:21 + 3 * j] = -Ji[:, 0:6].T
        A[6 + 3 * j:9 + 3 * j, 0:18] = Ji
        b[6 + 3 * j:9 + 3 * j] = -dyn["Jdv"][i] - prm["contact_damping"] * dyn["vf"][i]
    return A, b, cont


def tau_map(model, dyn, cont, n):
    """tau_k (internal joint order) = T z + t0."""
    T = np.zeros((12, n))
    T[:, 0:18] = dyn["M"][6:18, :]
    for j, i in enumerate(cont):
        T[3 * i:3 * i + 3, 18 + 3 * j:21 + 3 * j] = -dyn["L"][i].T
    return T, dyn["h"][6:18].copy()


def proto_step_id(model, prm, q, v, traj, contact):
    dyn = dynamics(model, q, v, gravity_in_bias=True)
    v_int = np.hstack([v[0:6], [v[model.v_index[k]] for k in range(12)]])
    A, b, cont = build_equalities(model, prm, dyn, v_int, contact)
    nc = len(cont)
    n = 18 + 3 * nc
    z0, Z, flag = nullspace(A, b)
    # task-space PD (inverse_dynamics_controller.py:187-197)
    rpy = rpy_from_R(dyn["R0"])
    N = rpy_N(rpy)
    rpyd = np.linalg.solve(N, v[0:3])
    p_nom, pd_nom, pdd_nom, rpy_nom, rpyd_nom, rpydd_nom = (traj[3 * i:3 * i + 3] for i in range(6))
    pdd_des = pdd_nom - prm["id_kp_body_p"] * (dyn["P"] - p_nom) - prm["id_kd_body_p"] * (v[3:6] - pd_nom)
    rpydd_des = rpydd_nom - prm["id_kp_body_rpy"] * (rpy - rpy_nom) - prm["id_kd_body_rpy"] * (rpyd - rpyd_nom)
    a_des = np.hstack([N @ rpydd_des, pdd_des])
    # cost rows: (weight, coefficient over z, target)
    rows = []
    for i in range(6):
        e = np.zeros(n); e[i] = 1.0
        rows.append((prm["id_w_body"], e, a_des[i]))
    err = np.sum((rpy - rpy_nom) ** 2) + np.sum((dyn["P"] - p_nom) ** 2)
    for i in range(4):
        if not contact[i]:
            Ji = foot_jacobian(dyn, i)
            pn, pdn, pddn = traj[18 + 3 * i:21 + 3 * i], traj[30 + 3 * i:33 + 3 * i], traj[42 + 3 * i:45 + 3 * i]
            a_s = pddn - prm["id_kp_foot"] * (dyn["p"][i] - pn) - prm["id_kd_foot"] * (dyn["vf"][i] - pdn)
            err += np.sum((dyn["p"][i] - pn) ** 2)
            for r in range(3):
                e = np.zeros(n); e[0:18] = Ji[r]
                rows.append((prm["id_w_foot"], e, a_s[r] - dyn["Jdv"][i][r]))
    for j in range(3 * nc):
        e = np.zeros(n); e[18 + j] = 1.0
        rows.append((prm["reg_f"], e, 0.0))
    T, t0 = tau_map(model, dyn, cont, n)
    if prm["reg_tau"] > 0:
        for k in range(12):
            rows.append((prm["reg_tau"], T[k], -t0[k]))
    if prm["reg_vd"] > 0:
        for k in range(18):
            e = np.zeros(n); e[k] = 1.0
            rows.append((prm["reg_vd"], e, 0.0))
    nf = Z.shape[1]
    Hr, gr = np.zeros((nf, nf)), np.zeros(nf)
    for w, e, d in rows:
        rz = e @ Z
        Hr += w * np.outer(rz, rz)
        gr += w * rz * (e @ z0 - d)
    # inequality rows
    G, hv = [], []
    mu = prm["mu"]
    for j in range(nc):
        for sx, sy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            e = np.zeros(n)
            e[18 + 3 * j], e[19 + 3 * j], e[20 + 3 * j] = sx, sy, -mu
            G.append(e @ Z); hv.append(-(e @ z0))
    if prm["torque_limits"]:
        for k in range(12):
            G.append(T[k] @ Z); hv.append(model.effort[k] - (T[k] @ z0 + t0[k]))
        for k in range(12):
            G.append(-T[k] @ Z); hv.append(model.effort[k] + (T[k] @ z0 + t0[k]))
    G = np.array(G).reshape(-1, nf)
    hv = np.array(hv)
    w, lam, iters, gflag = gi_solve(Hr, gr, G, hv)
    z = z0 + Z @ w
    tau_int = T @ z + t0
    tau = np.zeros(12)
    vd = np.zeros(18)
    vd[0:6] = z[0:6]
    for k in range(12):
        tau[model.act_index[k]] = tau_int[k]
        vd[model.v_index[k]] = z[6 + k]
    f = np.zeros((4, 3))
    for j, i in enumerate(cont):
        f[i] = z[18 + 3 * j:21 + 3 * j]
    return dict(tau=tau, vd=vd, f=f, iters=iters, flag=flag | (gflag << 1), err=err, lam=lam, cond=np.linalg.cond(Hr))


# ----------------------------------------------------------------------------- PC step
def proto_step_pc(model, prm, q, v, traj, contact):
    """Kernel formulation of pc_controller.py:43-255: tau_g cancels, delta is eliminated."""
    dyn = dynamics(model, q, v, gravity_in_bias=True)
    vidx = [model.v_index[k] for k in range(12)]
    to_int = lambda x: np.hstack([x[0:6], [x[i] for i in vidx]])            # noqa: E731  Drake -> internal order
    def to_drake(x):
        y = np.zeros(18); y[0:6] = x[0:6]
        for k in range(12):
            y[vidx[k]] = x[6 + k]
        return y
    v_int = to_int(v)
    bias = lambda vel_int: to_int_bias(model, q, to_drake(vel_int))         # noqa: E731
    A, b, cont = build_equalities(model, prm, dyn, v_int, contact)
    nc = len(cont)
    n = 18 + 3 * nc
    z0, Z, flag = nullspace(A, b)
    sw = [i for i in range(4) if not contact[i]]
    m = 6 + 3 * len(sw)
    M = dyn["M"]
    J = np.zeros((m, 18)); J[0:6, 0:6] = np.eye(6)
    for s_, i in enumerate(sw):
        J[6 + 3 * s_:9 + 3 * s_] = foot_jacobian(dyn, i)
    rpy = rpy_from_R(dyn["R0"]); N = rpy_N(rpy)
    xt = np.hstack([rpy - traj[9:12], dyn["P"] - traj[0:3]] + [dyn["p"][i] - traj[18 + 3 * i:21 + 3 * i] for i in sw])
    xdt = np.hstack([v[0:3] - N @ traj[12:15], v[3:6] - traj[3:6]] + [dyn["vf"][i] - traj[30 + 3 * i:33 + 3 * i] for i in sw])
    xddn = np.hstack([N @ traj[15:18], traj[6:9]] + [traj[42 + 3 * i:45 + 3 * i] for i in sw])
    kp = np.hstack([prm["pc_kp_body_rpy"] * np.ones(3), prm["pc_kp_body_p"] * np.ones(3), prm["pc_kp_foot"] * np.ones(3 * len(sw))])
    kd = np.hstack([prm["pc_kd_body_rpy"] * np.ones(3), prm["pc_kd_body_p"] * np.ones(3), prm["pc_kd_foot"] * np.ones(3 * len(sw))])
    wt = np.hstack([prm["pc_w_body"] * np.ones(6), prm["pc_w_foot"] * np.ones(3 * len(sw))])
    X = np.linalg.solve(M, J.T)
    Lam = np.linalg.inv(J @ X)
    s1 = Lam @ xdt
    w = v_int - X @ s1
    bv = bias(v_int)
    Cw = 0.5 * (bias(v_int + w) - bv - bias(w))
    Jdw = np.zeros(m)
    for s_, i in enumerate(sw):
        Jdw[6 + 3 * s_:9 + 3 * s_] = -skew(dyn["vf"][i] - v[3:6]) @ w[0:3] + dyn["Ld"][i] @ w[6 + 3 * i:9 + 3 * i]
    g0 = X.T @ (bv - Cw) - xddn + Jdw
    # reduced rows
    YJ = J @ Z[0:18, :]                 # task rows as functions of the reduced variables
    yJ0 = J @ z0[0:18]
    rows = []
    Wh = np.sqrt(wt)
    Rw = (Wh[:, None] * Lam) @ YJ
    r0 = Wh * (Lam @ (yJ0 + g0) + kp * xt + kd * xdt)
    nf = Z.shape[1]
    Hr = Rw.T @ Rw
    gr = Rw.T @ r0
    F = Z[18:18 + 3 * nc, :]
    Hr += prm["reg_f"] * F.T @ F
    gr += prm["reg_f"] * F.T @ z0[18:18 + 3 * nc]
    G, hv = [], []
    mu = prm["mu"]
    for j in range(nc):
        for sx, sy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            e = np.zeros(n); e[18 + 3 * j], e[19 + 3 * j], e[20 + 3 * j] = sx, sy, -mu
            G.append(e @ Z); hv.append(-(e @ z0))
    G.append(s1 @ YJ); hv.append(-(s1 @ (yJ0 + g0)) - xdt @ (kp * xt))
    G = np.array(G); hv = np.array(hv)
    wsol, lam, iters, gflag = gi_solve(Hr, gr, G, hv)
    z = z0 + Z @ wsol
    T, t0 = tau_map(model, dyn, cont, n)
    tau_int = T @ z + t0
    tau = np.zeros(12); vd = to_drake(z[0:18])
    for k in range(12):
        tau[model.act_index[k]] = tau_int[k]
    f = np.zeros((4, 3))
    for j, i in enumerate(cont):
        f[i] = z[18 + 3 * j:21 + 3 * j]
    V = 0.5 * xdt @ s1 + 0.5 * xt @ (kp * xt)
    Vdot = s1 @ (J @ z[0:18] + g0) + xdt @ (kp * xt)
    return dict(tau=tau, vd=vd, f=f, iters=iters, flag=flag | (gflag << 1), V=V, Vdot=Vdot, err=xt @ xt)


def to_int_bias(model, q, v_drake):
    d = dynamics(model, q, v_drake, gravity_in_bias=False)
    return d["h"]


# --------------------------------------------------------------- structured elimination (kernel v2)
def structured_nullspace(model, prm, dyn, contact):
    """z = z0 + Z w with w = [per leg: f_k (stance) or a_k (swing)] (12), using the block structure instead of a
    generic Gauss-Jordan: stance legs a_k = Li_k (r_k - Jb_k a_b); base rows give a 6x6 system for a_b."""
    cont = [i for i in range(4) if contact[i]]
    nc = len(cont)
    n = 18 + 3 * nc
    M, h = dyn["M"], dyn["h"]
    Sb = M[0:6, 0:6].copy()
    C = np.zeros((6, 12))
    c0 = -h[0:6].copy()
    AK = {}
    for k in range(4):
        Jb = np.hstack([-skew(dyn["rho"][k]), np.eye(3)])
        Mbk = M[0:6, 6 + 3 * k:9 + 3 * k]
        if contact[k]:
            Li = np.linalg.inv(dyn["L"][k])
            r = -dyn["Jdv"][k] - prm["contact_damping"] * dyn["vf"][k]
            AK[k] = (Li @ Jb, Li @ r)
            Sb -= Mbk @ AK[k][0]
            c0 -= Mbk @ AK[k][1]
            C[:, 3 * k:3 * k + 3] = Jb.T
        else:
            C[:, 3 * k:3 * k + 3] = -Mbk
    Bw = np.linalg.solve(Sb, C)          # a_b = Bw w + b0
    b0 = np.linalg.solve(Sb, c0)
    Z = np.zeros((n, 12)); z0 = np.zeros(n)
    Z[0:6] = Bw; z0[0:6] = b0
    for k in range(4):
        if contact[k]:
            s = cont.index(k)
            Z[6 + 3 * k:9 + 3 * k] = -AK[k][0] @ Bw
            z0[6 + 3 * k:9 + 3 * k] = AK[k][1] - AK[k][0] @ b0
            Z[18 + 3 * s:21 + 3 * s, 3 * k:3 * k + 3] = np.eye(3)
        else:
            Z[6 + 3 * k:9 + 3 * k, 3 * k:3 * k + 3] = np.eye(3)
    return z0, Z


def hybrid_nullspace(model, prm, dyn, contact):
    """Kernel v2 reduction: stance-leg joint accelerations analytically (a_k = Li_k (r_k - Jb_k a_b)), then pivoted
    Gauss-Jordan on the 6 base rows over u = [a_b(6); per leg: f_k (stance) or a_k (swing)] (18 columns)."""
    cont = [i for i in range(4) if contact[i]]
    nc = len(cont)
    n = 18 + 3 * nc
    M, h = dyn["M"], dyn["h"]
    B = np.zeros((6, 18)); c0 = -h[0:6].copy()
    B[:, 0:6] = M[0:6, 0:6]
    AK = {}
    for k in range(4):
        Jb = np.hstack([-skew(dyn["rho"][k]), np.eye(3)])
        Mbk = M[0:6, 6 + 3 * k:9 + 3 * k]
        if contact[k]:
            Li = np.linalg.inv(dyn["L"][k])
            r = -dyn["Jdv"][k] - prm["contact_damping"] * dyn["vf"][k]
            AK[k] = (Li @ Jb, Li @ r)
            B[:, 0:6] -= Mbk @ AK[k][0]
            c0 -= Mbk @ AK[k][1]
            B[:, 6 + 3 * k:9 + 3 * k] = -Jb.T
        else:
            B[:, 6 + 3 * k:9 + 3 * k] = Mbk
    u0, Zu, flag = nullspace(B, c0)          # u = u0 + Zu w, 12 free
    Z = np.zeros((n, 12)); z0 = np.zeros(n)
    Z[0:6] = Zu[0:6]; z0[0:6] = u0[0:6]
    for k in range(4):
        if contact[k]:
            s = cont.index(k)
            Z[6 + 3 * k:9 + 3 * k] = -AK[k][0] @ Zu[0:6]
            z0[6 + 3 * k:9 + 3 * k] = AK[k][1] - AK[k][0] @ u0[0:6]
            Z[18 + 3 * s:21 + 3 * s] = Zu[6 + 3 * k:9 + 3 * k]
            z0[18 + 3 * s:21 + 3 * s] = u0[6 + 3 * k:9 + 3 * k]
        else:
            Z[6 + 3 * k:9 + 3 * k] = Zu[6 + 3 * k:9 + 3 * k]
            z0[6 + 3 * k:9 + 3 * k] = u0[6 + 3 * k:9 + 3 * k]
    return z0, Z
