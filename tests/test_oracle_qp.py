"""Oracle QP fixtures and the KKT certificate of every golden solve."""
import numpy as np
import pytest

from oracle import controllers as oc
from oracle.qp import kkt_residuals, solve_qp

Q0 = np.array([1, 0, 0, 0, 0, 0, 0.3] + [0, -0.8, 1.6] * 4, float)
V0 = np.zeros(18)


def standing_at_fk(ctl):
    d = oc.standing_dict()
    c = ctl._common(Q0, V0, d)
    for i, f in enumerate(oc.FEET):
        d["p_" + f] = c.p_feet[i].copy()
    return d, c


def test_standing_fixture():
    """D.5: SimpleStanding with state = reference: vd ~ 0, sum f = m g, zero net moment, friction inactive."""
    ctl = oc.IDController("mini_cheetah", reg_f=1e-9)
    d, c = standing_at_fk(ctl)
    o = ctl.control_law(Q0, V0, d)
    assert o.status == "optimal" and o.active == []
    assert np.abs(o.vd).max() < 1e-6
    assert np.allclose(o.f.sum(0), [0, 0, 8.252 * 9.81], atol=1e-5)
    tg = ctl.plant.gravity_term(Q0)
    moment = sum(np.cross(c.p_feet[i] - Q0[4:7], o.f[i]) for i in range(4))
    assert np.allclose(moment, tg[:3], atol=1e-5)
    assert abs(o.objective) < 1e-9
    assert max(kkt_residuals(*o.qp, o.x, o.nu, o.lam)[:4]) < 1e-9


def test_edge_fixture_activates_friction():
    """D.6: EdgeTest (planners/simple.py:109-115) drives the friction rows active."""
    ctl = oc.IDController("mini_cheetah")
    d = oc.standing_dict()
    d["p_body"] = d["p_body"] + np.array([-0.1, 0.63, 0.0])
    o = ctl.control_law(Q0, V0, d)
    assert o.status in ("optimal", "ipm") and len(o.active) >= 4
    assert max(kkt_residuals(*o.qp, o.x, o.nu, o.lam)[:4]) < 1e-8
    mu = 0.7
    assert (np.abs(o.f[:, 0]) <= mu * o.f[:, 2] + 1e-8).all() and (np.abs(o.f[:, 1]) <= mu * o.f[:, 2] + 1e-8).all()


def test_raise_foot_fixture():
    """D.7: RaiseFoot (planners/simple.py:97-107): nc = 3 with one swing-foot cost."""
    ctl = oc.IDController("mini_cheetah")
    d = oc.standing_dict()
    d["p_body"] = d["p_body"] + np.array([-0.1, 0.05, 0.0])
    d["contact_states"] = [True, False, True, True]
    d["p_rf"] = d["p_rf"] + np.array([0, 0, 0.1])
    o = ctl.control_law(Q0, V0, d)
    assert o.x.size == 30 + 9 and np.all(o.f[1] == 0)
    assert max(kkt_residuals(*o.qp, o.x, o.nu, o.lam)[:4]) < 1e-8


def test_flight_has_no_force_rows():
    ctl = oc.IDController("mini_cheetah")
    d = oc.standing_dict()
    d["contact_states"] = [False] * 4
    o = ctl.control_law(Q0, V0 + 0.1, d)
    assert o.x.size == 30 and o.qp[4].shape[0] == 0     # inverse_dynamics_controller.py:216


def test_clf_care_closed_form():
    """per-channel CARE solution and gamma are constants."""
    ctl = oc.CLFController("mini_cheetah")
    d, _ = standing_at_fk(ctl)
    d["contact_states"] = [True, False, True, True]
    o = ctl.control_law(Q0, V0 + 0.05, d)
    P = o.P_lyap
    m = P.shape[0] // 2

    def care(qp, qd, r=1.0):
        p12 = np.sqrt(qp * r)
        p22 = np.sqrt(r * (qd + 2 * p12))
        return p12 * p22 / r, p12, p22
    b, f = care(5000, 200), care(200, 20)
    for i in range(m):
        p = b if i < 6 else f
        assert np.allclose([P[i, i], P[i, m + i], P[m + i, m + i]], p, rtol=1e-9)
    lam_max = max(np.linalg.eigvalsh(np.array([[b[0], b[1]], [b[1], b[2]]])).max(),
                  np.linalg.eigvalsh(np.array([[f[0], f[1]], [f[1], f[2]]])).max())
    assert np.isclose(o.gamma, 20.0 / lam_max, rtol=1e-9)
    assert np.isclose(lam_max, 1310.433, rtol=1e-6)


def test_qp_solver_small_known_answer():
    # min 1/2 |x|^2 - [1,1]x  s.t. x0 + x1 = 1, x0 <= 0.2  -> x = (0.2, 0.8)
    r = solve_qp(np.eye(2), -np.ones(2), np.array([[1.0, 1.0]]), np.array([1.0]), np.array([[1.0, 0.0]]), np.array([0.2]))
    assert np.allclose(r.x, [0.2, 0.8], atol=1e-10) and r.active == [0]


@pytest.mark.parametrize("case", ["cfg2_mini_cheetah_stand", "cfg3_anymal_trot", "cfg4_mini_cheetah_walk", "mixed_mini_cheetah"])
def test_oracle_reproduces_golden(case):
    """The committed golden vectors are what the oracle computes today (first 3 instances per case; all kinds)."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / f"{case}.npz")
    robot = "anymal_b" if "anymal" in case else "mini_cheetah"
    for kind, cls in (("id", oc.IDController), ("clf", oc.CLFController), ("pc", oc.PCController)):
        ctl = cls(robot)
        for i in range(3):
            if not g[f"{kind}_ok"][i]:
                continue
            o = ctl.control_law(g["q"][i], g["v"][i], oc.traj_to_dict(g["traj"][i], g["contact"][i]))
            assert np.abs(o.tau - g[f"{kind}_tau"][i]).max() < 1e-9
            assert max(kkt_residuals(*o.qp, o.x, o.nu, o.lam)[:4]) < 1e-7
            M = g["M"][i]
            assert np.abs(o.common.M - M).max() < 1e-12 * np.abs(M).max()
