"""Known answers for the rollout integrator of oracle/rollout.py (Drake-free; parity with Drake's own stepping is unpinned)."""
import numpy as np

from oracle import rollout as ro


def test_integrator_known_answers():
    q = np.array([1.0, 0, 0, 0, 0.1, -0.2, 0.3] + [0.0, -0.8, 1.6] * 4)
    v = np.zeros(18)
    # constant linear + joint velocity, zero acceleration: exact straight-line motion
    v[3:6] = [0.3, -0.1, 0.05]
    v[6:] = np.linspace(-1, 1, 12)
    qn, vn = q.copy(), v.copy()
    for _ in range(100):
        qn, vn = ro.integrate(qn, vn, np.zeros(18), 5e-3)
    assert np.allclose(qn[4:7], q[4:7] + 0.5 * v[3:6], atol=1e-14) and np.allclose(qn[7:], q[7:] + 0.5 * v[6:], atol=1e-13)
    assert np.array_equal(qn[:4], q[:4]) and np.array_equal(vn, v)
    # semi-implicit: the NEW velocity moves the position
    qn, vn = ro.integrate(q, np.zeros(18), np.r_[np.zeros(3), [2.0, 0, 0], np.zeros(12)], 0.01)
    assert np.isclose(vn[3], 0.02) and np.isclose(qn[4], q[4] + 0.01 * 0.02)
    # constant world-frame yaw rate: rotation about z by ~w t, unit norm kept
    w = 0.7
    qn, vn = q.copy(), np.zeros(18)
    vn[2] = w
    for _ in range(1000):
        qn, vn = ro.integrate(qn, vn, np.zeros(18), 1e-3)
    yaw = 2 * np.arctan2(qn[3], qn[0])
    assert abs(np.linalg.norm(qn[:4]) - 1) < 1e-15 and abs(yaw - w) < 1e-6 and abs(qn[1]) < 1e-15 and abs(qn[2]) < 1e-15
    # world-frame convention: R(q+) = (I + [w dt]x) R(q) to first order, i.e. the increment multiplies on the LEFT
    def rot(qt):
        w_, x, y, z = qt
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w_ * z), 2 * (x * z + w_ * y)],
                         [2 * (x * y + w_ * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w_ * x)],
                         [2 * (x * z - w_ * y), 2 * (y * z + w_ * x), 1 - 2 * (x * x + y * y)]])
    rng = np.random.default_rng(0)
    qq = rng.normal(size=4); qq /= np.linalg.norm(qq)
    om = rng.normal(size=3)
    vv = np.zeros(18); vv[:3] = om
    h = 1e-6
    qn, _ = ro.integrate(np.r_[qq, q[4:]], vv, np.zeros(18), h)
    W = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    assert np.abs(rot(qn[:4]) - (np.eye(3) + h * W) @ rot(qq)).max() < 1e-11


# ------------------------------------------------------------------------------------------ ground-contact plant
def _stand(plant_cls, robot="mini_cheetah"):
    from quadruped_drake_b200 import load_robot
    P = plant_cls(robot)
    q = load_robot(robot).nominal_q()
    pz = np.array([P.frame_position_quantities(q, np.zeros(18), f)[0][2] for f in P.foot_frames])
    q[6] -= pz.mean()                                    # feet exactly on the ground
    return P, q


def test_plant_standing_robot_is_carried_by_the_ground():
    """Known answers of the plant step: with the gravity-compensating joint torques of the standing posture the ground carries
    the weight (sum f_z = m g, 80.952 N), nothing moves, nothing penetrates; with zero torque the robot sags but the feet stay
    on the ground; in free fall (feet 0.2 m up) there is no contact force and v_z = -g dt."""
    from oracle import rollout as ro
    from oracle.dynamics import Plant
    P, q = _stand(Plant)
    v = np.zeros(18)
    # static equilibrium torques: base rows of tau_g = sum J' f with f = m g / 4 e_z per foot ... solve the 6 base rows for f_z only is
    # under-determined; use the ID controller's own answer instead (it stands, test_oracle_qp D.5)
    from oracle import controllers as oc
    d = oc.standing_dict()
    d["p_body"] = q[4:7].copy()
    for i, f in enumerate(P.foot_frames):
        d["p_" + oc.FEET[i]] = P.frame_position_quantities(q, v, f)[0]
    tau = oc.IDController(P).control_law(q, v, d).tau
    mg = P.total_mass() * 9.81
    qn, vn, f = ro.plant_step(P, q, v, tau, 5e-3, iters=400)       # converged Gauss-Seidel: statics up to the controller's own
    assert abs(f[:, 2].sum() - mg) < 1e-3 and (f[:, 2] > 0).all() and np.abs(vn).max() < 1e-6      # O(reg_f) residual acceleration
    qn, vn, f = ro.plant_step(P, q, v, tau, 5e-3)                  # the default 30 sweeps: within 1e-4 of it
    assert abs(f[:, 2].sum() - mg) < 1e-3 * mg and (f[:, 2] > 0).all()
    assert np.abs(vn).max() < 1e-4 and np.abs(qn - q).max() < 1e-6
    assert (np.abs(f[:, 0]) <= f[:, 2] + 1e-12).all() and (np.abs(f[:, 1]) <= f[:, 2] + 1e-12).all()
    # zero torque: the legs give way, the feet do not go through the floor
    qz, vz = q.copy(), v.copy()
    for _ in range(10):
        qz, vz, fz = ro.plant_step(P, qz, vz, np.zeros(12), 5e-3)
    pz = np.array([P.frame_position_quantities(qz, vz, fr)[0][2] for fr in P.foot_frames])
    assert pz.min() > -1e-3 and qz[6] < q[6] and fz[:, 2].sum() > 0.0
    # free fall
    qf = q.copy(); qf[6] += 0.2
    qn, vn, f = ro.plant_step(P, qf, v, np.zeros(12), 5e-3)
    assert np.abs(f).max() == 0.0 and abs(vn[5] + 9.81 * 5e-3) < 1e-12 and np.abs(np.delete(vn, 5)).max() < 1e-9


def test_plant_friction_cone_and_sliding():
    """A robot sliding sideways on its feet is decelerated by at most mu g; with mu = 0 it keeps sliding."""
    from oracle import rollout as ro
    from oracle.dynamics import Plant
    from oracle import controllers as oc
    P, q = _stand(Plant)
    v = np.zeros(18); v[3] = 1.0                                   # 1 m/s along x
    d = oc.standing_dict(); d["p_body"] = q[4:7].copy()
    tau = oc.IDController(P).control_law(q, np.zeros(18), d).tau
    for mu, expect in ((0.5, 0.5 * 9.81), (0.0, 0.0)):
        qn, vn, f = ro.plant_step(P, q, v, tau, 5e-3, mu=mu)
        assert (np.abs(f[:, 0]) <= mu * f[:, 2] + 1e-9).all()
        # every loaded foot slides: its friction force sits on the cone and opposes the motion (the normal forces themselves
        # shift with the pitching moment of the friction forces, so the deceleration is mu * sum f_z / m, not exactly mu g)
        assert np.abs(f[:, 0] + mu * f[:, 2]).max() < 1e-6 and f[:, 2].sum() > 0.5 * P.total_mass() * 9.81
        assert (vn[3] < 1.0 - 1e-4) if mu > 0 else (np.abs(f[:, :2]).max() == 0.0 and abs(vn[3] - 1.0) < 1e-3)


def test_emulated_plant_step_matches_oracle(built):
    """The device code of the plant step (host warp emulator) against the numpy restatement: standing, falling, sliding and
    random states, several steps."""
    import ctypes as C
    from pathlib import Path
    from oracle import rollout as ro
    from oracle.dynamics import Plant
    from quadruped_drake_b200 import load_robot
    from quadruped_drake_b200.capi import np_ptr
    emu = C.CDLL(str(Path(__file__).parent / "emu" / "libwbc_emu.so"))
    emu.emu_plant_step.argtypes = [C.c_void_p, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 5
    for robot in ("mini_cheetah", "anymal_b"):
        P, q0 = _stand(Plant, robot)
        ms = load_robot(robot).as_struct()
        rng = np.random.default_rng(3)
        n = 6
        q = np.tile(q0, (n, 1)); v = np.zeros((n, 18)); tau = rng.uniform(-3, 3, (n, 12))
        q[1, 6] += 0.1                                             # in the air
        v[2, 3:5] = [0.8, -0.4]                                    # sliding
        q[3:, 7:] += rng.uniform(-0.2, 0.2, (n - 3, 12)); v[3:] = rng.uniform(-1, 1, (n - 3, 18))
        q[4, 6] -= 0.004                                           # penetrating start
        qe, ve = q.copy(), v.copy()
        f, st = np.zeros((n, 4, 3)), np.zeros(n, np.int32)
        for step in range(3):
            ref = [ro.plant_step(P, qe[i] if step else q[i], ve[i] if step else v[i], tau[i], 5e-3, 1.0, 0.2, 30) for i in range(n)]
            assert emu.emu_plant_step(C.byref(ms), n, 5e-3, 1.0, 0.2, 30, np_ptr(qe), np_ptr(ve), np_ptr(tau), np_ptr(f), np_ptr(st)) == 0
            assert (st == 0).all()
            for i in range(n):
                assert np.abs(qe[i] - ref[i][0]).max() < 1e-9 and np.abs(ve[i] - ref[i][1]).max() < 1e-8, (robot, step, i)
                assert np.abs(f[i] - ref[i][2]).max() < 1e-6 * max(1.0, np.abs(ref[i][2]).max())
