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
