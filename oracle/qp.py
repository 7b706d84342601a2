"""ORACLE (test infrastructure only): exact dense convex QP solve.

Stands in for `OsqpSolver().Solve(mp)` (reference inverse_dynamics_controller.py:23,
223-225). OSQP (bundled with the reference's unpinned 2021-era Drake) is an ADMM
method run at eps_abs = eps_rel = 1e-3, far looser than the 1e-5
parity target, so parity is defined against the exact optimum of the same QP
(SURVEY.md 7 "Hard parts"). **parity unpinned**: the reference has no golden QP
solutions; this solver is pinned by its own KKT certificate, which every test checks, and
cross-checked by `oracle/osqp_admm.py`: OSQP's published ADMM iteration, run to convergence,
reaches the same point (tests/test_oracle_osqp.py).

    minimise 1/2 x'Px + q'x   s.t.  A x = b,  G x <= h

Method: Mehrotra predictor-corrector interior point on the full problem, then an
active-set polish (equality-constrained KKT solve on the identified active rows with
extended-precision iterative refinement, add/drop until the KKT signs hold).
"""
from __future__ import annotations

import numpy as np


class QPResult:
    def __init__(self, x, nu, lam, active, iters, status):
        self.x, self.nu, self.lam, self.active, self.iters, self.status = x, nu, lam, active, iters, status


def kkt_residuals(P, q, A, b, G, h, x, nu, lam):
    """(stationarity, equality, inequality violation, complementarity, min multiplier)."""
    r = P @ x + q
    if A.shape[0]:
        r = r + A.T @ nu
    if G.shape[0]:
        r = r + G.T @ lam
    eq = np.abs(A @ x - b).max() if A.shape[0] else 0.0
    s = h - G @ x if G.shape[0] else np.zeros(0)
    viol = max(0.0, (-s).max()) if s.size else 0.0
    comp = np.abs(lam * s).max() if s.size else 0.0
    lmin = lam.min() if lam.size else 0.0
    return np.abs(r).max(), eq, viol, comp, lmin


def _solve_sym(K, rhs):
    """Solve K y = rhs for a (possibly singular) symmetric KKT matrix, refined in long double."""
    if not (np.all(np.isfinite(K)) and np.all(np.isfinite(rhs))):
        raise np.linalg.LinAlgError("non-finite KKT system")
    try:
        y = np.linalg.solve(K, rhs)
        if not np.all(np.isfinite(y)):
            raise np.linalg.LinAlgError
    except np.linalg.LinAlgError:
        return np.linalg.lstsq(K, rhs, rcond=1e-13)[0]
    Kl, rl = K.astype(np.longdouble), rhs.astype(np.longdouble)
    for _ in range(3):
        res = (rl - Kl @ y.astype(np.longdouble)).astype(float)
        try:
            y = y + np.linalg.solve(K, res)
        except np.linalg.LinAlgError:
            break
    return y


def _ipm(P, q, A, b, G, h, max_iter=60, tol=1e-10):
    """Mehrotra predictor-corrector. Keeps the best iterate seen (smallest residual + gap) so that pushing the
    tolerance to round-off level can never return a worse or non-finite point."""
    n, me, mi = P.shape[0], A.shape[0], G.shape[0]
    x, nu = np.zeros(n), np.zeros(me)
    s, lam = np.ones(mi), np.ones(mi)
    if mi:
        s = np.maximum(h - G @ x, 1.0)
    best, best_score = (x, nu, s, lam), np.inf
    it = 0
    with np.errstate(all="ignore"):
        for it in range(1, max_iter + 1):
            rd = P @ x + q + (A.T @ nu if me else 0) + (G.T @ lam if mi else 0)
            re = A @ x - b if me else np.zeros(0)
            ri = G @ x + s - h if mi else np.zeros(0)
            mu = float(lam @ s) / mi if mi else 0.0
            score = max(np.abs(rd).max(), np.abs(re).max() if me else 0, np.abs(ri).max() if mi else 0, mu)
            if not np.isfinite(score):
                break
            if score < best_score:
                best, best_score = (x, nu, s, lam), score
            if score < tol:
                break
            d = lam / s if mi else np.zeros(0)
            H = P + (G.T * d) @ G if mi else P
            K = np.block([[H, A.T], [A, np.zeros((me, me))]]) if me else H
            if not np.all(np.isfinite(K)):
                break

            def step(rc):
                r1 = -rd + (G.T @ ((rc - lam * ri) / s) if mi else 0)  # eliminate ds, dlam
                rhs = np.hstack([r1, -re]) if me else r1
                sol = _solve_sym(K, rhs)
                dx, dnu = sol[:n], sol[n:]
                if mi:
                    ds = -ri - G @ dx
                    dlam = -(rc + lam * ds) / s
                else:
                    ds, dlam = np.zeros(0), np.zeros(0)
                return dx, dnu, ds, dlam

            def maxstep(z, dz):
                neg = dz < 0
                return min(1.0, float((-z[neg] / dz[neg]).min())) if neg.any() else 1.0

            try:
                if mi:
                    dx, dnu, ds, dlam = step(lam * s)
                    a = min(maxstep(s, ds), maxstep(lam, dlam))
                    mu_aff = float((lam + a * dlam) @ (s + a * ds)) / mi
                    sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
                    dx, dnu, ds, dlam = step(lam * s + ds * dlam - sigma * mu)
                    a = min(0.99 * min(maxstep(s, ds), maxstep(lam, dlam)), 1.0)
                    xn, nun, sn, lamn = x + a * dx, nu + a * dnu, s + a * ds, lam + a * dlam
                else:
                    dx, dnu, _, _ = step(np.zeros(0))
                    xn, nun, sn, lamn = x + dx, nu + dnu, s, lam
            except np.linalg.LinAlgError:
                break
            if not (np.all(np.isfinite(xn)) and np.all(np.isfinite(lamn)) and np.all(np.isfinite(sn))):
                break
            if mi and (sn.min() <= 0 or lamn.min() <= 0):
                break
            x, nu, s, lam = xn, nun, sn, lamn
    x, nu, s, lam = best
    return x, nu, s, lam, it


def _eq_solve(P, q, A, b, G, h, W):
    n, me = P.shape[0], A.shape[0]
    C = np.vstack([A, G[W]]) if len(W) else A
    d = np.hstack([b, h[W]]) if len(W) else b
    m = C.shape[0]
    K = np.block([[P, C.T], [C, np.zeros((m, m))]])
    sol = _solve_sym(K, np.hstack([-q, d]))
    return sol[:n], sol[n:n + me], sol[n + me:]


def _independent(A, G, W):
    """Greedy subset of W whose rows, together with A, are linearly independent."""
    keep = []
    base = A
    rank = np.linalg.matrix_rank(base) if base.shape[0] else 0
    for i in W:
        cand = np.vstack([base, G[i:i + 1]])
        r = np.linalg.matrix_rank(cand, tol=1e-9 * max(1.0, np.abs(cand).max()))
        if r > rank:
            keep.append(i)
            base, rank = cand, r
    return keep


def solve_qp(P, q, A, b, G, h, feas_tol=1e-9, max_polish=100):
    P, q = np.asarray(P, float), np.asarray(q, float)
    n = P.shape[0]
    A = np.zeros((0, n)) if A is None else np.asarray(A, float).reshape(-1, n)
    b = np.zeros(0) if b is None else np.asarray(b, float).ravel()
    G = np.zeros((0, n)) if G is None else np.asarray(G, float).reshape(-1, n)
    h = np.zeros(0) if h is None else np.asarray(h, float).ravel()
    mi = G.shape[0]
    x, nu, s, lam, it = _ipm(P, q, A, b, G, h, max_iter=80, tol=1e-12)
    status = "ipm"                      # interior-point answer kept unless the polish certifies a vertex
    if mi == 0:
        return QPResult(x, nu, lam, [], it, "optimal")
    W = [i for i in range(mi) if lam[i] > s[i]]
    gscale = np.maximum(1.0, np.abs(G).sum(axis=1))
    seen = set()
    for _ in range(max_polish):
        W = _independent(A, G, W)
        key = tuple(W)
        if key in seen:
            break                       # cycling on a degenerate vertex
        seen.add(key)
        xw, nuw, lw = _eq_solve(P, q, A, b, G, h, W)
        viol = (G @ xw - h) / gscale
        if len(W):
            viol[W] = 0.0
        worst = int(np.argmax(viol))
        if viol[worst] > feas_tol:
            W = sorted(W + [worst])
            continue
        if len(W) and lw.min() < -1e-9 * max(1.0, np.abs(lw).max()):
            W.pop(int(np.argmin(lw)))
            continue
        x, nu = xw, nuw
        lam = np.zeros(mi)
        if len(W):
            lam[W] = np.maximum(lw, 0.0)
        status = "optimal"
        break
    return QPResult(x, nu, lam, W, it, status)
