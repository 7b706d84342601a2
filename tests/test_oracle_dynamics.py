"""Known-answer tests pinning the oracle's Drake semantics."""
import numpy as np
import pytest

from oracle.dynamics import Plant, rpy_from_matrix, rpy_matrix, rpy_rate_matrix

Q0 = np.array([1, 0, 0, 0, 0, 0, 0.3] + [0, -0.8, 1.6] * 4, float)


def integrate(q, v, h):
    w = v[:3]
    ang = np.linalg.norm(w) * h
    ax = w / max(np.linalg.norm(w), 1e-300)
    dq = np.array([np.cos(ang / 2), *(np.sin(ang / 2) * ax)])
    w1, x1, y1, z1 = dq
    w2, x2, y2, z2 = q[:4] / np.linalg.norm(q[:4])
    out = q.copy()
    out[:4] = [w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
               w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2]
    out[4:7] += v[3:6] * h
    out[7:] += v[6:] * h
    return out


def random_state(plant, rng):
    q = Q0.copy() if plant.name == "mini_cheetah" else np.array([1, 0, 0, 0, 0, 0, .5] + [0, .4, -.8] * 2 + [0, -.4, .8] * 2, float)
    q[:4] = rng.normal(size=4)
    q[:4] /= np.linalg.norm(q[:4])
    q[4:7] += rng.uniform(-1, 1, 3)
    q[7:] += rng.uniform(-0.4, 0.4, 12)
    return q, rng.uniform(-1.5, 1.5, 18)


@pytest.mark.parametrize("robot,mass", [("mini_cheetah", 8.252), ("anymal_b", 30.421396462)])
def test_mass_and_gravity(robot, mass):
    P = Plant(robot)
    q, v = random_state(P, np.random.default_rng(3))
    M = P.mass_matrix(q)
    assert abs(P.total_mass() - mass) < 1e-9
    assert np.allclose(M, M.T, atol=1e-13)
    assert np.linalg.eigvalsh(M).min() > 0
    assert np.allclose(M[3:6, 3:6], mass * np.eye(3), atol=1e-12)
    tg = P.gravity_term(q)
    assert np.allclose(tg[3:6], [0, 0, mass * 9.81], atol=1e-10)       # +m g on z
    assert np.abs(P.bias_term(q, np.zeros(18))).max() == 0.0            # D.2: Cv(q, 0) = 0


def test_fk_fixture_mini_cheetah():
    """D.4: q0 of reference simulate.py:171-176 puts the LF foot at (0.17637, 0.111, -0.27798) in the body frame."""
    P = Plant("mini_cheetah")
    p, J, Jdv = P.frame_position_quantities(Q0, np.zeros(18), "LF_FOOT")
    assert np.allclose(p - Q0[4:7], [0.17637, 0.111, -0.27798], atol=2e-5)
    (R, pb), Jb, Jdvb = P.frame_pose_quantities(Q0, np.zeros(18), "body")
    assert np.allclose(Jb, np.hstack([np.eye(6), np.zeros((6, 12))]))   # A.1: J_body = [I6 0]
    assert np.allclose(Jdvb, 0)


@pytest.mark.parametrize("robot", ["mini_cheetah", "anymal_b"])
def test_jacobians_by_finite_differences(robot):
    """D.3: J = dp/dq N(q), Jdot v and Jdot by central differences along qdot = N(q) v."""
    P = Plant(robot)
    q, v = random_state(P, np.random.default_rng(5))
    h = 1e-6
    for f in P.foot_frames:
        p, J, Jdv = P.frame_position_quantities(q, v, f)
        pp, Jp, _ = P.frame_position_quantities(integrate(q, v, h), v, f)
        pm, Jm, _ = P.frame_position_quantities(integrate(q, v, -h), v, f)
        assert np.abs(J @ v - (pp - pm) / (2 * h)).max() < 1e-8
        assert np.abs(Jdv - (Jp - Jm) @ v / (2 * h)).max() < 1e-7
        assert np.abs(P.frame_jacobian_dot(q, v, f) - (Jp - Jm) / (2 * h)).max() < 1e-7


@pytest.mark.parametrize("robot", ["mini_cheetah", "anymal_b"])
def test_energy_identities(robot):
    """D.2: v'(Mdot - 2C)v = 0; C from polarisation reproduces the bias term."""
    P = Plant(robot)
    q, v = random_state(P, np.random.default_rng(6))
    h = 1e-6
    Md = (P.mass_matrix(integrate(q, v, h)) - P.mass_matrix(integrate(q, v, -h))) / (2 * h)
    Cv = P.bias_term(q, v)
    assert abs(v @ Md @ v - 2 * v @ Cv) < 1e-6
    C = P.coriolis_matrix(q, v)
    assert np.abs(C @ v - Cv).max() < 1e-12


def test_two_derivations_agree():
    """D.8: the world-frame composite formulation (tools/proto_kernel.py, what the CUDA kernel implements, on
    the welded/flattened model) against the projected Newton-Euler oracle on the raw tree."""
    from proto_kernel import dynamics, foot_jacobian
    from quadruped_drake_b200 import load_robot
    for robot in ["mini_cheetah", "anymal_b"]:
        P, model = Plant(robot), load_robot(robot)
        q, v = random_state(P, np.random.default_rng(8))
        M, Cv, tg, _ = P.calc_dynamics(q, v)
        d0 = dynamics(model, q, v, gravity_in_bias=False)
        d1 = dynamics(model, q, v, gravity_in_bias=True)
        assert np.abs(d0["M"] - M).max() < 1e-12 * np.abs(M).max()
        assert np.abs(d0["h"] - Cv).max() < 1e-12 * max(1, np.abs(Cv).max())
        assert np.abs(d1["h"] - Cv - tg).max() < 1e-12 * np.abs(tg).max()
        for k, f in enumerate(P.foot_frames):
            p, J, Jdv = P.frame_position_quantities(q, v, f)
            assert np.abs(J - foot_jacobian(d0, k)).max() < 1e-13
            assert np.abs(Jdv - d0["Jdv"][k]).max() < 1e-11


def test_breadth_first_order_is_a_permutation():
    Pd, Pb = Plant("mini_cheetah", "depth_first"), Plant("mini_cheetah", "breadth_first")
    rng = np.random.default_rng(9)
    q, v = random_state(Pd, rng)
    perm = [Pb.vidx[i] for i in range(len(Pd.names)) if Pd.vidx[i] is not None]   # depth position -> breadth index
    order = [Pd.vidx[i] for i in range(len(Pd.names)) if Pd.vidx[i] is not None]
    qb, vb = q.copy(), v.copy()
    for d, b in zip(order, perm):
        qb[b + 1], vb[b] = q[d + 1], v[d]
    Md, Mb = Pd.mass_matrix(q), Pb.mass_matrix(qb)
    idx = list(range(6)) + [0] * 12
    for d, b in zip(order, perm):
        idx[d] = b
    assert np.allclose(Md, Mb[np.ix_(idx, idx)], atol=1e-13)


def test_rpy_roundtrip():
    rng = np.random.default_rng(2)
    for _ in range(20):
        rpy = rng.uniform(-1.2, 1.2, 3)
        assert np.allclose(rpy_from_matrix(rpy_matrix(rpy)), rpy, atol=1e-12)
    N = rpy_rate_matrix([0.1, 0.2, 0.3])
    assert np.allclose(N[:, 2], [0, 0, 1])
