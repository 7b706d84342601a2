#!/usr/bin/env python
"""Generate tests/golden/*.npz: seeded inputs + oracle outputs for the dynamics and the three QP controllers.

The reference itself cannot run here (no pydrake/OSQP) and ships no golden data, so these vectors come
from oracle/ (parity unpinned - see oracle/dynamics.py). They freeze the oracle's answers so that (a) the
oracle cannot drift silently and (b) the GPU tests have fixed targets that do not need the slow Python
oracle at run time.   Usage: python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import oracle_fk  # noqa: E402
from oracle import controllers as oc  # noqa: E402
from oracle.dynamics import Plant  # noqa: E402
from quadruped_drake_b200 import load_robot  # noqa: E402
from quadruped_drake_b200.synth import generate  # noqa: E402

CASES = [  # name, robot, pattern, n, seed (seed = 20260117 + config index)
    ("cfg2_mini_cheetah_stand", "mini_cheetah", "stand", 24, 20260119),
    ("cfg3_anymal_trot", "anymal_b", "trot", 24, 20260120),
    ("cfg4_mini_cheetah_walk", "mini_cheetah", "walk", 24, 20260121),
    ("mixed_mini_cheetah", "mini_cheetah", "mixed", 32, 20260122),
]
CTRL = {"id": oc.IDController, "clf": oc.CLFController, "pc": oc.PCController, "mptc": oc.MPTCController}


def main(kinds):
    out = ROOT / "tests" / "golden"
    out.mkdir(exist_ok=True)
    for name, robot, pattern, n, seed in CASES:
        plant, model = Plant(robot), load_robot(robot)
        q, v, traj, contact = generate(model, n, seed, pattern, oracle_fk(plant))
        data = dict(q=q, v=v, traj=traj, contact=contact)
        M, Cv, tg = np.zeros((n, 18, 18)), np.zeros((n, 18)), np.zeros((n, 18))
        J, Jdv, pf = np.zeros((n, 4, 3, 18)), np.zeros((n, 4, 3)), np.zeros((n, 4, 3))
        for i in range(n):
            M[i], Cv[i], tg[i], _ = plant.calc_dynamics(q[i], v[i])
            for k, f in enumerate(plant.foot_frames):
                pf[i, k], J[i, k], Jdv[i, k] = plant.frame_position_quantities(q[i], v[i], f)
        data.update(M=M, Cv=Cv, tau_g=tg, J_feet=J, Jdv_feet=Jdv, p_feet=pf)
        for kind in kinds:
            ctl = CTRL[kind](plant)
            tau, vd, f, met, obj = np.zeros((n, 12)), np.zeros((n, 18)), np.zeros((n, 4, 3)), np.zeros((n, 4)), np.zeros(n)
            ok = np.zeros(n, bool)
            for i in range(n):
                if kind in ("pc", "mptc") and contact[i].sum() == 0:
                    continue                      # reference PC raises on full flight
                ctl.V = ctl.err = ctl.res = ctl.Vdot = 0.0
                o = ctl.control_law(q[i], v[i], oc.traj_to_dict(traj[i], contact[i]))
                tau[i], vd[i], f[i], met[i], obj[i] = o.tau, o.vd, o.f, o.metrics, o.objective
                ok[i] = o.status in ("optimal", "ipm")
            data.update({f"{kind}_tau": tau, f"{kind}_vd": vd, f"{kind}_f": f, f"{kind}_metrics": met,
                         f"{kind}_objective": obj, f"{kind}_ok": ok})
            print(name, kind, "solved", ok.sum(), "/", n)
        np.savez_compressed(out / f"{name}.npz", **data)


if __name__ == "__main__":
    main(sys.argv[1:] or ["id", "clf", "pc", "mptc"])
