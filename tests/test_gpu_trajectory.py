"""GPU parity of the trajectory sampler (wbc_plan_create / wbc_sample_trajectory) against oracle/trajectory.py, which
reproduces the reference's compiled TOWR spline code bit for bit (tests/test_oracle_trajectory.py).

Tolerance: the device evaluates the cubic in Horner form with FMAs and takes the local time as t - (running sum), while
the reference sums pow(t, c) * coeff and subtracts the durations one by one (polynomial.cc:49-63, spline.cc:68-79), so
values agree to rounding, not bit for bit: |diff| <= 1e-12 * max(1, |ref|) (positions / velocities O(1), accelerations
O(100)). Segment selection, contact flags and the nearest-sample lookup are integer decisions and must agree exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def close(a, b):
    return np.all(np.abs(a - b) <= TOL * np.maximum(1.0, np.abs(b)))


@pytest.fixture(scope="module")
def ctl(built):
    from quadruped_drake_b200.controller import BatchedController
    c = BatchedController("mini_cheetah", device=0)
    yield c
    c.close()


def oracle_plan(robot, combo, **kw):
    from oracle import trajectory as tr
    return tr.make_gait_plan(robot, combo, 5.0, kw.get("goal", (1.5, 0.0)), kw.get("swing_height", 0.05), kw.get("yaw_goal", 0.0))


@pytest.mark.parametrize("robot,combo", [("mini_cheetah", 0), ("anymal_b", 1), ("mini_cheetah", 2), ("mini_cheetah", 3), ("anymal_b", 4)])
def test_continuous_sampling_matches_oracle(ctl, robot, combo):
    from quadruped_drake_b200 import planner as pl
    plan = pl.make_gait_plan(robot, combo, 5.0, (1.5, 0.2), 0.06, 0.3)
    ora = oracle_plan(robot, combo, goal=(1.5, 0.2), swing_height=0.06, yaw_goal=0.3)
    s = pl.TrajectorySampler(ctl, plan)
    rng = np.random.default_rng(combo)
    junctions = np.concatenate([np.cumsum(ora.phase_dur[ee])[:-1] for ee in range(4)] + [np.cumsum(ora.base_linear.durations)[:-1]])
    t = np.concatenate([rng.uniform(0, 5, 400), junctions, junctions + 1e-11, junctions - 1e-11, junctions + 2e-10, [0.0, 5.0]])
    o = s.sample(t, forces=True)
    assert (o["status"] == 0).all() and np.array_equal(o["t_eval"], t)
    for i, ti in enumerate(t):
        traj, contact, f = ora.sample(float(ti))
        assert close(o["traj"][i], traj), (ti, np.abs(o["traj"][i] - traj).max())
        assert np.array_equal(o["contact"][i], contact), ti
        assert close(o["f"][i], f)


def test_planner_grid_mode_matches_towr_planner(ctl):
    """planners/towr.py:92-148: SimpleStanding before wait_time, then the nearest of the 5001 stored samples."""
    from oracle import trajectory as tr
    from quadruped_drake_b200 import planner as pl
    planner = pl.TowrTrunkPlanner(ctl, robot="mini_cheetah", gait="walk")
    ora = oracle_plan("mini_cheetah", 0)
    ts = tr.publish_timestamps(5.0)
    assert np.array_equal(planner.plan.grid, ts)
    rng = np.random.default_rng(1)
    t = np.concatenate([rng.uniform(0, 6, 500), [0.0, 0.999999, 1.0, 1.0005, 1.0004999, 6.0], 1.0 + 0.5 * (ts[100:110] + ts[101:111])])
    o = planner.sample(t, forces=True)
    for i, ti in enumerate(t):
        traj, contact, f = tr.towr_planner_output(ora, ts, float(ti))
        if ti >= 1.0:
            assert o["t_eval"][i] == ts[np.abs(ts - (ti - 1.0)).argmin()], ti          # exact sample choice
        else:
            assert o["t_eval"][i] == -1.0
        assert close(o["traj"][i], traj) and np.array_equal(o["contact"][i], contact) and close(o["f"][i], f), ti
    d = planner.SetTrunkOutputs(0.3)           # the reference dict (planners/simple.py:45-85)
    assert d["contact_states"] == [True] * 4 and np.array_equal(d["p_body"], [0.0, 0.0, 0.3]) and d["f_cj"].shape == (3, 4)
    d = planner.SetTrunkOutputs(2.5)
    assert set(d) >= {"p_lf", "pd_rh", "pdd_lh", "rpy_body", "pdd_body", "contact_states", "f_cj", "u2_max"}


def test_multiple_plans_status_and_device_path(ctl):
    import torch
    from quadruped_drake_b200 import planner as pl
    plans = [pl.make_gait_plan("mini_cheetah", c) for c in range(5)]
    oras = [oracle_plan("mini_cheetah", c) for c in range(5)]
    s = pl.TrajectorySampler(ctl, plans)
    rng = np.random.default_rng(2)
    n = 3000
    t, pi = rng.uniform(0, 5, n), rng.integers(0, 5, n).astype(np.int32)
    t[:4] = [-0.5, 5.5, 1.0, 2.0]
    pi[2], pi[3] = 7, -1
    o = s.sample(torch.from_numpy(t).cuda(), torch.from_numpy(pi).cuda(), forces=False)
    torch.cuda.synchronize()
    st = o["status"].cpu().numpy()
    assert st[0] == 1 and st[1] == 1 and st[2] == 2 and st[3] == 2 and (st[4:] == 0).all()
    traj, contact = o["traj"].cpu().numpy(), o["contact"].cpu().numpy()
    assert close(traj[0], oras[pi[0]].sample(0.0)[0]) and close(traj[1], oras[pi[1]].sample(5.0)[0])
    for i in range(4, 300):
        r, c, _ = oras[pi[i]].sample(float(t[i]))
        assert close(traj[i], r) and np.array_equal(contact[i], c)


def test_full_size_config3_properties_and_closed_chain(ctl):
    """BASELINE configs[2] shape: anymal_b trot, 16384 instances. Size-independent properties of the sampled reference,
    then the whole chain sampler -> ID-QP step on the device with every instance solved."""
    import torch
    from quadruped_drake_b200 import planner as pl
    from quadruped_drake_b200.controller import BatchedController
    n = 16384
    plan = pl.make_gait_plan("anymal_b", "trot", 5.0, (1.5, 0.0), 0.08)
    s = pl.TrajectorySampler(ctl, plan)
    rng = np.random.default_rng(3)
    t = np.sort(rng.uniform(0, 5, n))
    tt = torch.from_numpy(t).cuda()
    o = s.sample(tt, forces=True)
    h = 1e-6
    op, om = s.sample(tt + h), s.sample(tt - h)
    traj, contact, f = o["traj"].cpu().numpy(), o["contact"].cpu().numpy(), o["f"].cpu().numpy().reshape(n, 4, 3)
    tp, tm = op["traj"].cpu().numpy(), om["traj"].cpu().numpy()
    inside = (t > 1e-3) & (t < 5 - 1e-3)
    # velocity is the derivative of position, acceleration of velocity (central differences; away from junction kinks of acc)
    assert np.abs((tp[inside, :3] - tm[inside, :3]) / (2 * h) - traj[inside, 3:6]).max() < 1e-6
    assert np.abs((tp[inside, 18:30] - tm[inside, 18:30]) / (2 * h) - traj[inside, 30:42]).max() < 1e-5
    feet_p, feet_v = traj[:, 18:30].reshape(n, 4, 3), traj[:, 30:42].reshape(n, 4, 3)
    assert (feet_p[..., 2][contact == 1] == 0).all() and (feet_v[contact == 1] == 0).all()      # stance feet rest
    assert (feet_p[..., 2] >= -1e-12).all() and (feet_p[..., 2] <= 0.08 + 1e-9).all()
    assert (f[contact == 0] == 0).all() and (f[..., 2] >= -1e-9).all()
    pat = set(map(tuple, np.unique(contact, axis=0)))
    assert pat <= {(1, 1, 1, 1), (1, 0, 0, 1), (0, 1, 1, 0), (0, 0, 0, 0)}                        # stand, bP, Pb, flight
    # chain: robot standing at the nominal configuration tracks the sampled reference
    anymal = BatchedController("anymal_b", device=0, torque_limits=1)
    from quadruped_drake_b200.synth import nominal_state
    q, v = nominal_state(anymal.model, n)
    q[:, 4:7] = traj[:, 0:3]
    out = anymal.step("id", torch.from_numpy(q).cuda(), torch.from_numpy(v).cuda(), o["traj"], o["contact"])
    torch.cuda.synchronize()
    st = out.status.cpu().numpy()
    assert (st == 0).all(), np.unique(st, return_counts=True)
    assert np.isfinite(out.tau.cpu().numpy()).all()


def test_step_with_device_plan_equals_sample_then_step(ctl):
    """wbc_step_plan_host: the host sends q, v, t; the trajectory rows are sampled on the device. Same torques as sampling
    first and stepping with host trajectory buffers - for pageable arrays (staged) and page-locked ones (zero-copy)."""
    from quadruped_drake_b200 import capi
    from quadruped_drake_b200 import planner as pl
    from quadruped_drake_b200.synth import generate
    n = 4099                                                    # page-locked buffers: two halves on two streams, ragged tail
    q, v, _, _ = generate(ctl.model, n, 5, "stand", ctl.fk)
    q[:, 4:6] = 0.0
    bh = float(ctl.model.nominal_q()[6])
    s = pl.TrajectorySampler(ctl, [pl.make_motion_plan("mini_cheetah", m, 6.0, base_height=bh) for m in ("standing", "raise_foot")])
    rng = np.random.default_rng(1)
    t, pi = rng.uniform(0, 6, n), rng.integers(0, 2, n).astype(np.int32)
    o = s.sample(t, pi)
    ref = ctl.step("id", q, v, o["traj"], o["contact"])
    a = ctl.step_plan("id", s, q, v, t, pi)
    assert np.array_equal(a.tau, ref.tau) and np.array_equal(a.status, ref.status) and np.array_equal(a.metrics, ref.metrics)
    hq, hv, ht, hpi = capi.pinned_empty((n, 19)), capi.pinned_empty((n, 18)), capi.pinned_empty((n,)), capi.pinned_empty((n,), np.int32)
    hq[:], hv[:], ht[:], hpi[:] = q, v, t, pi
    tau, met, st = capi.pinned_empty((n, 12)), capi.pinned_empty((n, 4)), capi.pinned_empty((n,), np.int32)
    b = ctl.step_plan("id", s, hq, hv, ht, hpi, tau, met, st)
    assert np.array_equal(b.tau, ref.tau) and np.array_equal(b.status, ref.status)
    c = ctl.step_plan("clf", s, q, v, t)                       # no plan index: plan 0 for everyone
    o0 = s.sample(t)
    assert np.array_equal(c.tau, ctl.step("clf", q, v, o0["traj"], o0["contact"]).tau)


def test_sampler_kernel_shapes_agree(ctl):
    """wbc_sample_trajectory picks the instances-per-warp shape by batch size (2 below 8192 samples, 8 below 32768, 32 above;
    csrc/wbc_traj.cuh): the same (plan, t) pairs through all three give identical bits, ragged tails, bad plan indices, clamped
    times, grid-mode plans and planned forces included."""
    from quadruped_drake_b200 import planner as pl
    plans = [pl.make_gait_plan("mini_cheetah", c) for c in range(4)]
    s = pl.TrajectorySampler(ctl, plans)
    rng = np.random.default_rng(11)
    n = 40003
    t, pi = rng.uniform(-0.2, 5.3, n), rng.integers(-1, 5, n).astype(np.int32)
    big = s.sample(t, pi, forces=True)                                   # lane = instance
    for step in (3001, 10007):                                           # 16 lanes / 4 lanes per instance
        for o in range(0, n, step):
            part = s.sample(t[o:o + step], pi[o:o + step], forces=True)
            for key in ("traj", "contact", "f", "status", "t_eval"):
                assert np.array_equal(np.asarray(part[key]), np.asarray(big[key])[o:o + step]), (step, o, key)
    grid = pl.TowrTrunkPlanner(ctl, robot="mini_cheetah", gait="trot")
    tg = rng.uniform(0.0, 6.0, 33000)
    bigg = grid.sample(tg, forces=True)
    for o in (0, 9000, 29000):
        part = grid.sample(tg[o:o + 4000], forces=True)
        for key in ("traj", "contact", "f", "t_eval"):
            assert np.array_equal(np.asarray(part[key]), np.asarray(bigg[key])[o:o + 4000]), (o, key)
