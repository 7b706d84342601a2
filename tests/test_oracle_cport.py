"""The C twin of the oracle (oracle/c/oracle_id.c, the timed CPU baseline) against the numpy oracle's golden vectors.
Its interior-point QP has no active-set polish, so only the well-conditioned quantities (dynamics, vd, net contact
wrench) are compared tightly; tau/f carry the IPM's error along the tie-break directions (documented in its header)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).parent / "golden"


@pytest.mark.parametrize("case", ["cfg2_mini_cheetah_stand", "cfg3_anymal_trot", "mixed_mini_cheetah"])
def test_cport_matches_numpy_oracle(built, case):
    from oracle.cport import LIB, id_batch
    from quadruped_drake_b200 import load_robot
    g = np.load(GOLD / f"{case}.npz")
    robot = "anymal_b" if "anymal" in case else "mini_cheetah"
    tau, vd, f, st = id_batch(robot, g["q"], g["v"], g["traj"], g["contact"], threads=2)
    assert (st == 0).all()
    assert np.abs(vd - g["id_vd"]).max() < 1e-5
    # tau / f: the IPM stops at a 1e-11 KKT score, which leaves O(1e-2) along the reg_f = 1e-6 tie-break directions for a few
    # of the 256 instances; the port is the timed baseline, never the checker
    err = np.abs(tau - g["id_tau"]).max(axis=1)
    assert err.max() < 0.25 and np.median(err) < 1e-3 and np.abs(f.sum(1) - g["id_f"].sum(1)).max() < 1e-4
    # dynamics of the port itself: bit-level agreement with the numpy oracle
    lib = C.CDLL(str(LIB))
    ms = load_robot(robot).as_struct()
    P = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    for i in range(3):
        q, v = np.ascontiguousarray(g["q"][i]), np.ascontiguousarray(g["v"][i])
        M, Cv, tg = np.zeros((18, 18)), np.zeros(18), np.zeros(18)
        J, Jdv, p = np.zeros((4, 3, 18)), np.zeros((4, 3)), np.zeros((4, 3))
        lib.oracle_dynamics(C.byref(ms), P(q), P(v), P(M), P(Cv), P(tg), P(J), P(Jdv), P(p))
        for name, a in (("M", M), ("Cv", Cv), ("tau_g", tg), ("J_feet", J), ("Jdv_feet", Jdv), ("p_feet", p)):
            assert np.abs(a - g[name][i]).max() < 1e-12 * max(1.0, np.abs(g[name][i]).max()), name
