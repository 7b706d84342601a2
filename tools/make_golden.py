#!/usr/bin/env python
"""Generate tests/golden/*.npz: seeded inputs + oracle outputs for the dynamics and the QP controllers.

The reference itself cannot run here (no pydrake/OSQP) and ships no golden data, so these vectors come
from oracle/ (parity unpinned - see oracle/dynamics.py). They freeze the oracle's answers so that (a) the
oracle cannot drift silently and (b) the GPU tests have fixed targets that do not need the slow Python
oracle at run time. If pydrake is importable (`tools/pydrake_probe.py`), `--drake` adds the reference
classes' own answers under `drake_*` keys; the tests prefer those when present.

Every case holds >= 256 instances (a process pool runs the oracle; ~0.3 s per instance and controller);
the dense dynamics terms (M, J) are kept for the first N_DYN instances only, to keep the fixtures small.

Usage: python tools/make_golden.py [case-name-substring ...] [--jobs J]
"""
import argparse
import os
import sys
from multiprocessing import get_context
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N_DYN = 32
ALL = ("id", "clf", "pc", "mptc")
QP3 = ("id", "clf", "pc")
# name, robot, pattern, n, seed (seed = 20260117 + config index), dof_order, oracle params, kinds
CASES = [
    ("cfg2_mini_cheetah_stand", "mini_cheetah", "stand", 256, 20260119, "depth_first", {}, ALL),
    ("cfg3_anymal_trot", "anymal_b", "trot", 256, 20260120, "depth_first", {}, ALL),
    ("cfg4_mini_cheetah_walk", "mini_cheetah", "walk", 256, 20260121, "depth_first", {}, ALL),
    ("mixed_mini_cheetah", "mini_cheetah", "mixed", 256, 20260122, "depth_first", {}, ALL),
    # 2021-era Drake velocity numbering: all abductions, all hips, all knees
    ("bf_mini_cheetah_mixed", "mini_cheetah", "mixed", 256, 20260123, "breadth_first", {}, ALL),
    ("bf_anymal_trot", "anymal_b", "trot", 256, 20260124, "breadth_first", {}, ALL),
    # optional |tau| <= effort box (BASELINE configs[2]); states chosen so that the box is active for many instances
    ("tl_mini_cheetah_walk", "mini_cheetah", "walk", 256, 20260125, "depth_first", {"torque_limits": 1}, ALL),
    ("tl_anymal_trot", "anymal_b", "trot+", 256, 20260126, "depth_first", {"torque_limits": 1}, ALL),
    # the reference's manual test motions (planners/simple.py:87-115) as direct step cases
    ("fixtures_mini_cheetah", "mini_cheetah", "fixtures", 288, 20260127, "depth_first", {}, ALL),
]


def fixture_inputs(model, n, seed, fk):
    """States around the standing posture with the trunk targets of OrientationTest / RaiseFoot / EdgeTest
    (reference planners/simple.py:87-115); instance 3 j + k uses fixture k at time t_j."""
    from oracle.controllers import dict_to_traj, standing_dict
    rng = np.random.default_rng(seed)
    qn = model.nominal_q()
    q = np.tile(qn, (n, 1))
    v = np.zeros((n, 18))
    # first three instances: exactly the nominal state at rest (simulate.py:171-179); the others are perturbed
    q[3:, 7:] += rng.uniform(-0.15, 0.15, (n - 3, 12))
    q[3:, 4:7] += rng.uniform(-0.03, 0.03, (n - 3, 3))
    ang = rng.uniform(-0.15, 0.15, (n - 3, 3))
    from quadruped_drake_b200.synth import rpy_to_quat
    q[3:, 0:4] = rpy_to_quat(ang)
    v[3:, 0:6] = rng.uniform(-0.2, 0.2, (n - 3, 6))
    v[3:, 6:] = rng.uniform(-1.0, 1.0, (n - 3, 12))
    traj, contact = np.zeros((n, 54)), np.zeros((n, 4), np.uint8)
    for i in range(n):
        k, t = i % 3, 6.0 * (i // 3) / (n // 3)                      # t sweeps the reference's 6 s (simulate.py:21)
        d = standing_dict("mini_cheetah")
        if k == 0:                                                   # OrientationTest(t)
            d["rpy_body"] = np.array([0.0, 0.4 * np.sin(t), 0.4 * np.cos(t)])
            d["rpyd_body"] = np.array([0.0, 0.4 * np.cos(t), -0.4 * np.sin(t)])
            d["rpydd_body"] = np.array([0.0, -0.4 * np.sin(t), -0.4 * np.cos(t)])
        elif k == 1:                                                 # RaiseFoot(t)
            d["p_body"] = d["p_body"] + np.array([-0.1, 0.05, 0.0])
            if t > 1:
                d["contact_states"] = [True, False, True, True]
                d["p_rf"] = d["p_rf"] + np.array([0.0, 0.0, 0.1])
        else:                                                        # EdgeTest()
            d["p_body"] = d["p_body"] + np.array([-0.1, 0.63, 0.0])
        traj[i], contact[i] = dict_to_traj(d)
    return q, v, traj, contact


def make_inputs(robot, pattern, n, seed, dof_order):
    from conftest import oracle_fk
    from oracle.dynamics import Plant
    from quadruped_drake_b200 import load_robot
    from quadruped_drake_b200.synth import generate
    plant, model = Plant(robot, dof_order), load_robot(robot, dof_order=dof_order)
    fk = oracle_fk(plant)
    if pattern == "fixtures":
        return fixture_inputs(model, n, seed, fk)
    if pattern.endswith("+"):
        # larger tracking errors (x2.5 on the position noise) so that the torque box of the heavier robot becomes active
        q, v, traj, contact = generate(model, n, seed, pattern[:-1], fk)
        p_feet, _ = fk(q, v)
        traj[:, 0:3] = q[:, 4:7] + 2.5 * (traj[:, 0:3] - q[:, 4:7])
        traj[:, 18:30] = (p_feet + 2.5 * (traj[:, 18:30].reshape(n, 4, 3) - p_feet)).reshape(n, 12)
        return q, v, traj, contact
    return generate(model, n, seed, pattern, fk)


def _chunk(args):
    robot, dof_order, params, kinds, q, v, traj, contact, lo, n_dyn = args
    from oracle import controllers as oc
    from oracle.dynamics import Plant
    ctrl = {"id": oc.IDController, "clf": oc.CLFController, "pc": oc.PCController, "mptc": oc.MPTCController}
    plant = Plant(robot, dof_order)
    n = len(q)
    out = {}
    nd = max(0, min(n, n_dyn - lo))
    M, Cv, tg = np.zeros((nd, 18, 18)), np.zeros((n, 18)), np.zeros((n, 18))
    J, Jdv, pf = np.zeros((nd, 4, 3, 18)), np.zeros((n, 4, 3)), np.zeros((n, 4, 3))
    for i in range(n):
        Mi, Cv[i], tg[i], _ = plant.calc_dynamics(q[i], v[i])
        if i < nd:
            M[i] = Mi
        for k, f in enumerate(plant.foot_frames):
            pf[i, k], Ji, Jdv[i, k] = plant.frame_position_quantities(q[i], v[i], f)
            if i < nd:
                J[i, k] = Ji
    out.update(M=M, Cv=Cv, tau_g=tg, J_feet=J, Jdv_feet=Jdv, p_feet=pf)
    for kind in kinds:
        ctl = ctrl[kind](plant, **params)
        tau, vd, f, met, obj = np.zeros((n, 12)), np.zeros((n, 18)), np.zeros((n, 4, 3)), np.zeros((n, 4)), np.zeros(n)
        ok = np.zeros(n, bool)
        for i in range(n):
            if kind in ("pc", "mptc") and contact[i].sum() == 0:
                continue                      # reference PC raises on full flight
            ctl.V = ctl.err = ctl.res = ctl.Vdot = 0.0
            try:
                o = ctl.control_law(q[i], v[i], oc.traj_to_dict(traj[i], contact[i]))
            except Exception:  # noqa: BLE001  (infeasible torque box etc.: recorded as not ok)
                continue
            tau[i], vd[i], f[i], met[i], obj[i] = o.tau, o.vd, o.f, o.metrics, o.objective
            ok[i] = o.status in ("optimal", "ipm") and o.primal_res < 1e-7
        out.update({f"{kind}_tau": tau, f"{kind}_vd": vd, f"{kind}_f": f, f"{kind}_metrics": met,
                    f"{kind}_objective": obj, f"{kind}_ok": ok})
    return lo, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(exist_ok=True)
    for name, robot, pattern, n, seed, dof_order, params, kinds in CASES:
        if args.names and not any(s in name for s in args.names):
            continue
        q, v, traj, contact = make_inputs(robot, pattern, n, seed, dof_order)
        per = max(4, (n + 4 * args.jobs - 1) // (4 * args.jobs))
        jobs = [(robot, dof_order, params, kinds, q[lo:lo + per], v[lo:lo + per], traj[lo:lo + per], contact[lo:lo + per], lo, N_DYN)
                for lo in range(0, n, per)]
        with get_context("spawn").Pool(args.jobs) as pool:
            parts = sorted(pool.map(_chunk, jobs), key=lambda r: r[0])
        data = dict(q=q, v=v, traj=traj, contact=contact, dof_order=np.array(dof_order),
                    params=np.array(repr(sorted(params.items()))), n_dyn=np.array(N_DYN))
        for key in parts[0][1]:
            data[key] = np.concatenate([p[1][key] for p in parts], axis=0)
        for kind in kinds:
            print(name, kind, "solved", int(data[f"{kind}_ok"].sum()), "/", n, flush=True)
        np.savez_compressed(out_dir / f"{name}.npz", **data)


if __name__ == "__main__":
    main()
