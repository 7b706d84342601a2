"""ORACLE (test infrastructure only): restatement of the solver the reference actually calls.

`OsqpSolver().Solve(mp)` (reference inverse_dynamics_controller.py:23,223-225; clf_controller.py:228; pc_controller.py:249;
mptc_controller.py:304) hands the QP to OSQP, which lives in pydrake - an un-vendored, unpinned dependency of the reference
(mid/late-2021 Drake bundling OSQP 0.6.x) that is absent from /root/reference and not installable here. This file
restates OSQP's PUBLISHED algorithm (Stellato, Banjac, Goulart, Bemporad, Boyd: "OSQP: an operator splitting solver for quadratic
programs", Math. Prog. Comp. 12, 2020: Algorithm 1, sections 3.4 termination, 5.1 preconditioning, 5.2 rho selection, 4 polishing)
with the library's default settings and Drake's one override (`polish = 1`), in dense numpy:

    minimise 1/2 x'Px + q'x   s.t.  l <= A x <= u

**parity unpinned**: there is no OSQP binary here to compare iterates with, and OSQP's own adaptive-rho schedule is timing
dependent by default (this restatement uses the fixed interval OSQP falls back to when profiling is off). The file exists to answer
one question with the reference's own algorithm class: how far is "what OSQP returns" from the exact optimum that
`oracle/qp.py` computes and the CUDA path is compared with? (tests/test_oracle_osqp.py) It is not on any product
path and not used by the GPU parity tests.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

OSQP_INFTY = 1e30
MIN_SCALING, MAX_SCALING = 1e-4, 1e4
RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_EQ_OVER_RHO_INEQ = 1e3
RHO_TOL = 1e-4


@dataclass
class Settings:
    """OSQP 0.6 defaults (osqp/include/constants.h); `polish` is Drake's default (OsqpSolver turns it on)."""
    rho: float = 0.1
    sigma: float = 1e-6
    alpha: float = 1.6
    eps_abs: float = 1e-3
    eps_rel: float = 1e-3
    max_iter: int = 4000
    check_termination: int = 25
    scaling: int = 10
    adaptive_rho: bool = True
    adaptive_rho_interval: int = 100          # OSQP's fixed fallback (ADAPTIVE_RHO_FIXED); the timed default is not reproducible
    adaptive_rho_tolerance: float = 5.0
    polish: bool = True
    delta: float = 1e-6
    polish_refine_iter: int = 3


@dataclass
class Result:
    x: np.ndarray
    y: np.ndarray
    status: str            # "solved" | "max_iter"
    iters: int
    rho_updates: int
    pri_res: float
    dua_res: float
    polished: bool         # the polished point replaced the ADMM point
    x_admm: np.ndarray     # the un-polished ADMM iterate (what OSQP returns with polish = 0)


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


def _ruiz(P, q, A, l, u, iters):
    """Modified Ruiz equilibration of the KKT matrix + cost scaling (paper 5.1, Algorithm 2)."""
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, q, A = P.copy(), q.copy(), A.copy()
    for _ in range(iters):
        dn = np.maximum(np.abs(P).max(axis=0) if n else 0.0, np.abs(A).max(axis=0) if m else np.zeros(n))
        en = np.abs(A).max(axis=1) if m else np.zeros(0)
        Dt, Et = 1.0 / np.sqrt(_limit(dn)), 1.0 / np.sqrt(_limit(en))
        P = Dt[:, None] * P * Dt[None, :]
        A = Et[:, None] * A * Dt[None, :]
        q = Dt * q
        D, E = D * Dt, E * Et
        pn = np.abs(P).max(axis=0).mean()
        qn = np.abs(q).max()
        ct = 1.0 / _limit(np.array([max(pn, qn)]))[0]
        P, q, c = P * ct, q * ct, c * ct
    return P, q, A, l * E, u * E, D, E, c


def _rho_vec(l, u, rho):
    eq = (u - l) < RHO_TOL
    loose = (l < -OSQP_INFTY * MIN_SCALING) & (u > OSQP_INFTY * MIN_SCALING)
    r = np.where(eq, RHO_EQ_OVER_RHO_INEQ * rho, rho)
    return np.where(loose, RHO_MIN, r)


def _residuals(P0, q0, A0, x, z, y):
    """Unscaled residuals and the normalisers of the termination test (paper 3.4)."""
    Ax, Px, Aty = A0 @ x, P0 @ x, A0.T @ y
    inf = lambda v: float(np.abs(v).max()) if v.size else 0.0  # noqa: E731
    return inf(Ax - z), inf(Px + q0 + Aty), max(inf(Ax), inf(z)), max(inf(Px), inf(Aty), inf(q0))


def solve(P, q, A, l, u, settings: Settings | None = None) -> Result:
    s = settings or Settings()
    P0, q0, A0 = np.asarray(P, float), np.asarray(q, float), np.asarray(A, float)
    l0 = np.maximum(np.asarray(l, float), -OSQP_INFTY)
    u0 = np.minimum(np.asarray(u, float), OSQP_INFTY)
    n, m = P0.shape[0], A0.shape[0]
    Ps, qs, As, ls, us, D, E, c = _ruiz(P0, q0, A0, l0, u0, s.scaling) if s.scaling else (P0, q0, A0, l0, u0, np.ones(n), np.ones(m), 1.0)
    rho = s.rho
    rv = _rho_vec(ls, us, rho)

    def factor(rv):
        K = np.block([[Ps + s.sigma * np.eye(n), As.T], [As, -np.diag(1.0 / rv)]])
        return np.linalg.inv(K)          # dense stand-in for the quasi-definite LDL' of OSQP (QDLDL)

    Kinv = factor(rv)
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)          # cold start: the reference sets no initial guess
    status, it, rho_updates = "max_iter", 0, 0
    pri = dua = np.inf
    for it in range(1, s.max_iter + 1):
        sol = Kinv @ np.hstack([s.sigma * x - qs, z - y / rv])
        xt, nu = sol[:n], sol[n:]
        zt = z + (nu - y) / rv
        x = s.alpha * xt + (1.0 - s.alpha) * x
        zr = s.alpha * zt + (1.0 - s.alpha) * z
        z_new = np.clip(zr + y / rv, ls, us)
        y = y + rv * (zr - z_new)
        z = z_new
        check = s.check_termination and it % s.check_termination == 0
        adapt = s.adaptive_rho and s.adaptive_rho_interval and it % s.adaptive_rho_interval == 0
        if check or adapt or it == s.max_iter:
            xu, zu, yu = D * x, z / E, E * y / c
            pri, dua, npri, ndua = _residuals(P0, q0, A0, xu, zu, yu)
            if check and pri <= s.eps_abs + s.eps_rel * npri and dua <= s.eps_abs + s.eps_rel * ndua:
                status = "solved"
                break
            if adapt:
                # paper 5.2: rho <- rho sqrt( (r_prim / max(|Ax|, |z|)) / (r_dual / max(|Px|, |A'y|, |q|)) ), in the SCALED space
                Ax, Px, Aty = As @ x, Ps @ x, As.T @ y
                inf = lambda v: float(np.abs(v).max()) if v.size else 0.0  # noqa: E731
                p_s, d_s = inf(Ax - z), inf(Px + qs + Aty)
                pn, dn = max(inf(Ax), inf(z)), max(inf(Px), inf(Aty), inf(qs))
                est = np.sqrt((p_s / (pn + 1e-10)) / (d_s / (dn + 1e-10) + 1e-300)) * rho if d_s > 0 else rho
                est = min(max(est, RHO_MIN), RHO_MAX)
                if est > rho * s.adaptive_rho_tolerance or est < rho / s.adaptive_rho_tolerance:
                    rho, rho_updates = est, rho_updates + 1
                    rv = _rho_vec(ls, us, rho)
                    Kinv = factor(rv)
    xu, zu, yu = D * x, z / E, E * y / c
    pri, dua, _, _ = _residuals(P0, q0, A0, xu, zu, yu)
    x_admm = xu.copy()
    polished = False
    if s.polish and status == "solved":
        # paper 4: active rows guessed from the scaled iterate, then the equality-constrained KKT system with delta regularisation
        # and iterative refinement (polish.c)
        low = (z - ls) < -y
        upp = (us - z) < y
        act = np.where(low | upp)[0]
        Ared = As[act]
        rhs = np.hstack([-qs, np.where(low[act], ls[act], us[act])])
        k = len(act)
        K = np.block([[Ps, Ared.T], [Ared, np.zeros((k, k))]])
        Kreg = K + np.diag(np.hstack([np.full(n, s.delta), np.full(k, -s.delta)]))
        Kri = np.linalg.inv(Kreg)
        t = Kri @ rhs
        for _ in range(s.polish_refine_iter):
            t = t + Kri @ (rhs - K @ t)
        xp = t[:n]
        yp = np.zeros(m)
        yp[act] = t[n:]
        zp = np.clip(As @ xp, ls, us)
        xpu, zpu, ypu = D * xp, zp / E, E * yp / c
        ppri, pdua, _, _ = _residuals(P0, q0, A0, xpu, zpu, ypu)
        if (ppri < pri and pdua < dua) or (ppri < pri and dua < 1e-10) or (pdua < dua and pri < 1e-10):
            xu, yu, pri, dua, polished = xpu, ypu, ppri, pdua, True
    return Result(xu, yu, status, it, rho_updates, pri, dua, polished, x_admm)


def solve_reference_qp(P, q, A, b, G, h, settings: Settings | None = None) -> Result:
    """The controllers' QP in the row order Drake's OsqpSolver builds: linear equalities (l = u), then inequalities (l = -inf)."""
    Aq = np.vstack([A, G]) if G.shape[0] else A
    lo = np.hstack([b, np.full(G.shape[0], -np.inf)])
    up = np.hstack([b, h])
    return solve(P, q, Aq, lo, up, settings)
