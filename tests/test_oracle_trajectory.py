"""oracle/trajectory.py pinned three ways: (1) against the reference's own TOWR sources compiled into
oracle/_ref/libtowr_ref.so (skipped where /root/reference was never available), (2) against the golden vectors that
library produced (tests/golden/towr_spline.npz, tools/make_golden_traj.py), (3) Hermite known answers. Also checks
the product's host-side plan builder (quadruped_drake_b200/planner.py) against the oracle's restatement."""
from pathlib import Path

import numpy as np
import pytest

from oracle import trajectory as tr
from oracle import towr_ref

G = np.load(Path(__file__).parent / "golden" / "towr_spline.npz")
needs_ref = pytest.mark.skipif(not towr_ref.available(), reason="oracle/_ref/libtowr_ref.so not built (needs /root/reference)")


def test_gait_tables_match_golden():
    for combo in range(5):
        for ee in range(4):
            d = np.array(tr.phase_durations(combo, 5.0, ee))
            assert np.array_equal(d, G[f"phase_c{combo}_e{ee}"]), (combo, ee)
            assert tr.contact_at_start(combo, ee) == bool(G["contact_start"][combo, ee])


def test_spline_points_match_golden_bit_exactly():
    for k in range(int(G["n_splines"])):
        d, nodes, ts, ref = G[f"dur_{k}"], G[f"nodes_{k}"], G[f"ts_{k}"], G[f"pts_{k}"]
        s = tr.Spline(d, nodes)
        for t, r, seg in zip(ts, ref, G[f"seg_{k}"]):
            assert tr.segment_id(t, list(d)) == seg
            assert np.array_equal(np.concatenate(s.point(t)), r)


def test_contact_flags_match_golden():
    for combo in range(5):
        for ee in range(4):
            pd = tr.phase_durations(combo, 5.0, ee)
            c0 = tr.contact_at_start(combo, ee)
            got = [tr.is_contact_phase(t, pd, c0) for t in G["contact_ts"]]
            assert got == [bool(x) for x in G["contact_flags"][combo, ee]]


@needs_ref
def test_live_reference_library_agrees():
    rng = np.random.default_rng(7)
    for combo in range(5):
        for ee in range(4):
            assert np.array_equal(np.array(tr.phase_durations(combo, 3.7, ee)), towr_ref.phase_durations(combo, 3.7, ee))
    for _ in range(20):
        n = int(rng.integers(1, 30))
        d, nodes = rng.uniform(0.03, 0.5, n), rng.normal(0, 2, (n + 1, 6))
        ts = np.concatenate([rng.uniform(0, d.sum() * 0.999, 64), np.cumsum(d)[:-1]])
        ref = towr_ref.spline_points(d, nodes, ts)
        s = tr.Spline(d, nodes)
        for t, r in zip(ts, ref):
            assert np.array_equal(np.concatenate(s.point(t)), r)
            assert tr.segment_id(t, list(d)) == towr_ref.segment_id(t, d)


def test_hermite_known_answers():
    # interpolation conditions and a cubic reproduced exactly
    p0, v0, p1, v1, T = np.array([1.0, -2.0, 0.5]), np.array([0.3, 0.0, -1.0]), np.array([2.0, 1.0, 0.0]), np.array([-0.5, 2.0, 0.25]), 0.7
    c = tr.hermite_coeff(p0, v0, p1, v1, T)
    p, v, _ = tr.poly_point(c, 0.0)
    assert np.allclose(p, p0, atol=1e-15) and np.allclose(v, v0, atol=1e-15)
    p, v, _ = tr.poly_point(c, T)
    assert np.allclose(p, p1, atol=1e-13) and np.allclose(v, v1, atol=1e-13)
    f = lambda t: 1 + 2 * t - 3 * t ** 2 + 0.5 * t ** 3           # noqa: E731
    fd = lambda t: 2 - 6 * t + 1.5 * t ** 2                         # noqa: E731
    c = tr.hermite_coeff([f(0)] * 3, [fd(0)] * 3, [f(T)] * 3, [fd(T)] * 3, T)
    p, v, a = tr.poly_point(c, 0.31)
    assert abs(p[0] - f(0.31)) < 1e-13 and abs(v[0] - fd(0.31)) < 1e-13 and abs(a[0] - (-6 + 3 * 0.31)) < 1e-12


def test_publish_timestamps_and_planner_lookup():
    ts = tr.publish_timestamps(5.0)
    assert len(ts) == 5001 and ts[-1] == 5.0 and ts[0] == 0.0       # planners/towr.py stores ~5001 messages
    assert ts[1000] != 1.0                                            # accumulated floating-point time, not k * dt
    plan = tr.make_gait_plan("mini_cheetah", 0)
    traj, contact, f = tr.towr_planner_output(plan, ts, 0.5)
    assert np.array_equal(traj, tr.simple_standing_traj()[0]) and contact.all()
    traj, contact, f = tr.towr_planner_output(plan, ts, 1.0 + 1.23449)
    t_near = ts[np.abs(ts - 1.23449).argmin()]
    assert np.array_equal(traj, plan.sample(t_near)[0])


def test_synthetic_plan_is_consistent():
    for robot, combo in (("mini_cheetah", 0), ("anymal_b", 1)):
        plan = tr.make_gait_plan(robot, combo)
        for t in np.linspace(0, 5, 101):
            traj, contact, f = plan.sample(float(t))
            feet_z, feet_v = traj[18:30].reshape(4, 3)[:, 2], traj[30:42].reshape(4, 3)
            assert (feet_z[contact == 1] == 0).all() and np.abs(feet_v[contact == 1]).max(initial=0) == 0   # stance feet rest on the ground
            assert (feet_z >= -1e-12).all() and (f.reshape(4, 3)[contact == 0] == 0).all()


def test_product_plan_builder_matches_oracle():
    """quadruped_drake_b200/planner.py (host-side setup, separate implementation) builds the same tables."""
    from quadruped_drake_b200 import planner as pl
    for robot, combo in (("mini_cheetah", 0), ("anymal_b", 1), ("mini_cheetah", 4)):
        a, b = tr.make_gait_plan(robot, combo, 5.0, (1.5, 0.2), 0.06, 0.3), pl.make_gait_plan(robot, combo, 5.0, (1.5, 0.2), 0.06, 0.3, sample_dt=1e-3)
        for sa, sb in [(a.base_linear, b.base_linear), (a.base_angular, b.base_angular)] + list(zip(a.ee_motion, b.ee_motion)) + list(zip(a.ee_force, b.ee_force)):
            assert np.array_equal(np.array(sa.durations), sb.durations) and np.array_equal(sa.nodes, sb.nodes)
        for ee in range(4):
            assert np.array_equal(np.array(a.phase_dur[ee]), b.phase_durations[ee]) and a.contact_start[ee] == b.contact_at_start[ee]
        assert np.array_equal(b.grid, tr.publish_timestamps(5.0))
    t, c = pl.simple_standing("mini_cheetah")
    assert np.array_equal(t, tr.simple_standing_traj()[0])
