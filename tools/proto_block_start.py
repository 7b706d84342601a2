#!/usr/bin/env python
"""Numpy prototype for VERDICT r1 item 4 ("cut the pivots"): would starting the dual active-set method from the block of
initially violated rows save pivots? Uses the oracle's QPs (reduced to the null space of the equalities - the pivot sequence of
Goldfarb-Idnani is basis independent), a dense reference implementation of the method with the kernel's pivot rule (most violated
row), and a block start: greedily independent violated rows, constrained minimiser on them, rows with a negative multiplier dropped
one at a time, then the ordinary loop. Test infrastructure / experiment only (imports oracle/).

  python tools/proto_block_start.py mini_cheetah stand 200 20260119

Measured (200 / 150 instances): configs[1] stand: 10.1 pivots (1.6 drops, final active set 6.9 - the kernel reports 9.84 / 6.63)
against 5.8 block adds + 0.8 block drops + 4.8 pivots; walk: 6.0 against 3.9 + 0.7 + 2.7. A block add skips the pivot selection,
the ratio test and the step (about half a pivot), so the loop work falls by ~16 % (stand) / ~10 % (walk) - 7 % of the step at
best, for a second code path in a kernel that already stalls on instruction fetch. Not built."""
import sys
from pathlib import Path

import numpy as np
import scipy.linalg as sla

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import oracle_fk  # noqa: E402
from oracle import controllers as oc  # noqa: E402
from oracle.dynamics import Plant  # noqa: E402
from quadruped_drake_b200 import load_robot  # noqa: E402
from quadruped_drake_b200.synth import generate  # noqa: E402


def reduced(P, q, A, b, G, h):
    Z = sla.null_space(A)
    x0 = np.linalg.lstsq(A, b, rcond=None)[0]
    return Z.T @ P @ Z, Z.T @ (P @ x0 + q), G @ Z, h - G @ x0


def gi(H, g, C, d, A0=None, x0=None, u0=None):
    """Goldfarb-Idnani with the most-violated pivot rule -> (iterations, drops, active set, x, ok)."""
    n = len(g)
    Hi = np.linalg.inv(H)
    A, x, u = ([], -Hi @ g, np.zeros(0)) if A0 is None else (list(A0), x0.copy(), u0.copy())
    iters = drops = 0
    while True:
        viol = -(d - C @ x)
        viol[A] = -1
        p = int(np.argmax(viol)) if len(viol) else 0
        if not len(viol) or viol[p] < 1e-9 * (1 + abs(d[p])):
            return iters, drops, A, x, True
        up = 0.0
        while True:
            iters += 1
            if iters > 200:
                return iters, drops, A, x, False
            if A:
                N = C[A].T
                K = np.block([[H, N], [N.T, np.zeros((len(A), len(A)))]])
                sol = np.linalg.solve(K, np.hstack([C[p], np.zeros(len(A))]))
                z, r = sol[:n], sol[n:]
            else:
                z, r = Hi @ C[p], np.zeros(0)
            nz = C[p] @ z
            t2 = -(d[p] - C[p] @ x) / nz if nz > 1e-12 else np.inf
            t1, l = np.inf, -1
            for k in range(len(A)):
                if r[k] > 1e-14 and u[k] / r[k] < t1:
                    t1, l = u[k] / r[k], k
            t = min(t1, t2)
            if not np.isfinite(t):
                return iters, drops, A, x, False
            if np.isfinite(t2):
                x = x - t * z
            if len(A):
                u = u - t * r
            up += t
            if t2 <= t1:
                A.append(p)
                u = np.append(u, up)
                break
            drops += 1
            A.pop(l)
            u = np.delete(u, l)


def block_start(H, g, C, d):
    n = len(g)
    x = -np.linalg.solve(H, g)
    s = d - C @ x
    L = np.linalg.cholesky(H)
    V = []
    for i in np.argsort(s):
        if s[i] >= -1e-9 * (1 + abs(d[i])):
            break
        cand = V + [int(i)]
        Nw = np.linalg.solve(L, C[cand].T)
        if np.linalg.svd(Nw, compute_uv=False)[-1] > 1e-7 * np.linalg.norm(Nw[:, -1]):
            V = cand
    nadd, ndrop, xV, u = len(V), 0, x, np.zeros(0)
    while V:
        N = C[V].T
        K = np.block([[H, N], [N.T, np.zeros((len(V), len(V)))]])
        sol = np.linalg.solve(K, np.hstack([-g, d[V]]))
        xV, u = sol[:n], sol[n:]
        if (u >= -1e-12).all():
            break
        V.pop(int(np.argmin(u)))
        ndrop += 1
    return V, xV, u, nadd, ndrop


if __name__ == "__main__":
    robot, pattern, n, seed = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    plant, model = Plant(robot), load_robot(robot)
    q, v, traj, contact = generate(model, n, seed, pattern, oracle_fk(plant))
    ctl = oc.IDController(plant)
    rows = []
    for i in range(n):
        o = ctl.control_law(q[i], v[i], oc.traj_to_dict(traj[i], contact[i]))
        if not o.qp[4].shape[0]:
            continue
        H, g, C, d = reduced(*o.qp)
        it, dr, Aset, x, ok = gi(H, g, C, d)
        V, xV, u, nadd, nd = block_start(H, g, C, d)
        it2, _, _, x2, ok2 = gi(H, g, C, d, V, xV, u) if V else gi(H, g, C, d)
        assert ok and ok2 and np.abs(x2 - x).max() < 1e-6 * (1 + np.abs(x).max())
        rows.append((it, dr, len(Aset), nadd, nd, it2))
    r = np.array(rows, float)
    print(f"{robot} {pattern}: one-at-a-time {r[:, 0].mean():.2f} pivots ({r[:, 1].mean():.2f} drops, final active set {r[:, 2].mean():.2f}); "
          f"block start {r[:, 3].mean():.2f} adds + {r[:, 4].mean():.2f} drops + {r[:, 5].mean():.2f} pivots; "
          f"loop work at half a pivot per block add: {(0.5 * r[:, 3] + r[:, 4] + r[:, 5]).mean():.2f} vs {r[:, 0].mean():.2f}")
