"""GPU parity of the closed-loop rollout (wbc_rollout / wbc_integrate) against oracle/rollout.py.

Tolerance: every step adds the controller's own tolerance (vd within 1e-6 of the oracle's exact QP optimum) times dt, and the
closed loop is contracting (PD task gains), so after K steps states agree to ~1e-7; the test asserts 1e-6 on q and 1e-5 on v
and the last torques (the bar BASELINE.json sets for QP torques)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
Q0 = np.array([1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.3] + [0.0, -0.8, 1.6] * 4)      # simulate.py:171-176


@pytest.fixture(scope="module")
def ctl(built):
    from quadruped_drake_b200.controller import BatchedController
    c = BatchedController("mini_cheetah", device=0)
    yield c
    c.close()


def test_integrate_matches_oracle(ctl):
    import torch
    from oracle import rollout as ro
    from quadruped_drake_b200.rollout import integrate
    rng = np.random.default_rng(0)
    n = 1000
    q = rng.normal(size=(n, 19)); q[:, :4] /= np.linalg.norm(q[:, :4], axis=1, keepdims=True)
    v, vd = rng.normal(size=(n, 18)), rng.normal(0, 10, (n, 18))
    tq, tv, tvd, tt = (torch.from_numpy(x.copy()).cuda() for x in (q, v, vd, np.zeros(n)))
    integrate(ctl, tq, tv, tvd, 5e-3, tt)
    torch.cuda.synchronize()
    for i in range(0, n, 37):
        qn, vn = ro.integrate(q[i], v[i], vd[i], 5e-3)
        assert np.abs(tq[i].cpu().numpy() - qn).max() < 1e-15 and np.abs(tv[i].cpu().numpy() - vn).max() < 1e-15
    assert np.allclose(tt.cpu().numpy(), 5e-3)


@pytest.mark.parametrize("kind,combo,steps", [("id", 0, 24), ("clf", 1, 16), ("pc", 0, 8)])
def test_rollout_matches_oracle(ctl, kind, combo, steps):
    from oracle import rollout as ro, trajectory as tr
    from quadruped_drake_b200 import planner as pl
    from quadruped_drake_b200.rollout import rollout
    plan = pl.make_gait_plan("mini_cheetah", combo)
    ora = tr.make_gait_plan("mini_cheetah", combo)
    s = pl.TrajectorySampler(ctl, plan)
    t0 = np.array([0.0, 0.29, 0.9, 2.0])            # standing, just before the first lift-off, in the gait, mid motion
    n = len(t0)
    q0 = np.tile(Q0, (n, 1))
    for i, t in enumerate(t0):                       # start on the reference (base at the planned position)
        q0[i, 4:7] = ora.sample(float(t))[0][0:3]
    r = rollout(ctl, s, kind, q0, np.zeros((n, 18)), t0, steps, 5e-3, log_metrics=True)
    assert (r.status_or == 0).all()
    for i in range(n):
        q, v, t, tau, log = ro.rollout("mini_cheetah", kind, ora, q0[i], np.zeros(18), float(t0[i]), steps, 5e-3)
        assert np.abs(r.q[i] - q).max() < 1e-6, (i, np.abs(r.q[i] - q).max())
        assert np.abs(r.v[i] - v).max() < 1e-5
        assert np.abs(r.tau[i] - tau).max() < 1e-5
        assert abs(r.t[i] - t) < 1e-12
        assert np.abs(r.metrics_log[:, i, 1] - log[:, 1]).max() < 1e-8      # tracking error of every step
        assert np.isclose(r.err_max[i], log[:, 1].max(), atol=1e-8)


def test_graph_replay_equals_plain_launches_and_device_path(ctl):
    import torch
    from quadruped_drake_b200 import planner as pl
    from quadruped_drake_b200.rollout import rollout
    s = pl.TrajectorySampler(ctl, pl.make_gait_plan("mini_cheetah", "trot"))
    n, steps = 512, 40
    rng = np.random.default_rng(1)
    t0 = rng.uniform(0, 4.5, n)
    q0 = np.tile(Q0, (n, 1))
    q0[:, 4:7] = s.sample(t0)["traj"][:, 0:3]          # start on the planned base position
    outs = []
    for graph in (False, True):
        q, v, t = (torch.from_numpy(x).cuda() for x in (q0.copy(), np.zeros((n, 18)), t0.copy()))
        with torch.cuda.stream(torch.cuda.Stream()):
            r = rollout(ctl, s, "id", q, v, t, steps, 5e-3, use_graph=graph, log_metrics=True)
        torch.cuda.synchronize()
        outs.append([x.cpu().numpy() for x in (r.q, r.v, r.t, r.tau, r.metrics_log, r.status_or)])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert (outs[0][5] == 0).all() and np.allclose(outs[0][2], t0 + steps * 5e-3)


def quat_to_rpy(q):
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return np.stack([np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y)), np.arcsin(np.clip(2 * (w * y - z * x), -1, 1)),
                     np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))], axis=1)


def test_full_size_reference_test_motions_six_seconds(ctl):
    """The reference run itself, batched: 4096 robots x 1200 steps (simulate.py:20-21: dt 5e-3, sim_time 6.0) from
    simulate.py's q0, tracking the manual test motions of planners/simple.py:87-115 at 24 different phases, closed loop on
    the device. Size-independent properties: every QP solved, unit quaternions, the base converges to the reference."""
    import torch
    from quadruped_drake_b200 import planner as pl
    from quadruped_drake_b200.rollout import rollout
    bh = Q0[6] - ctl.dynamics(Q0[None], np.zeros((1, 18)))["p_feet"][0, :, 2].mean()      # feet exactly on the ground
    plans = [pl.make_motion_plan("mini_cheetah", m, 6.0, base_height=bh, phase=ph)
             for m in ("orientation", "heave", "raise_foot") for ph in np.linspace(0, 2 * np.pi, 8, endpoint=False)]
    s = pl.TrajectorySampler(ctl, plans)
    n, steps, dt = 4096, 1200, 5e-3
    rng = np.random.default_rng(2)
    pi = rng.integers(0, len(plans), n).astype(np.int32)
    q0 = np.tile(Q0, (n, 1)); q0[:, 6] = bh
    q0[:, 7:] += rng.uniform(-0.05, 0.05, (n, 12))
    q, v, t, tpi = (torch.from_numpy(x).cuda() for x in (q0, np.zeros((n, 18)), np.zeros(n), pi))
    with torch.cuda.stream(torch.cuda.Stream()):
        r = rollout(ctl, s, "id", q, v, t, steps, dt, plan_index=tpi)
    torch.cuda.synchronize()
    st = r.status_or.cpu().numpy()
    assert (st == 0).all(), np.unique(st, return_counts=True)
    qf, tf = q.cpu().numpy(), t.cpu().numpy()
    assert np.allclose(tf, 6.0) and np.abs(np.linalg.norm(qf[:, :4], axis=1) - 1).max() < 1e-14
    ref = s.sample(np.full(n, 6.0 - 1e-9), pi)["traj"]
    assert np.abs(qf[:, 4:7] - ref[:, 0:3]).max() < 5e-3, np.abs(qf[:, 4:7] - ref[:, 0:3]).max()
    assert np.abs(quat_to_rpy(qf) - ref[:, 9:12]).max() < 5e-3
    assert r.metrics.cpu().numpy()[:, 1].max() < 1e-3         # tracking-error metric of the last step


def test_gait_start_and_frozen_failures(ctl):
    """Walk plan through the first lift-offs (all solved), then a dynamically inconsistent long trot: instances that fail or
    diverge are frozen and flagged instead of spreading NaNs (the reference would assert, inverse_dynamics_controller.py:224)."""
    import torch
    from quadruped_drake_b200 import planner as pl
    from quadruped_drake_b200.rollout import rollout
    bh = Q0[6] - ctl.dynamics(Q0[None], np.zeros((1, 18)))["p_feet"][0, :, 2].mean()
    n = 4096
    rng = np.random.default_rng(5)
    q0 = np.tile(Q0, (n, 1)); q0[:, 6] = bh
    q0[:, 7:] += rng.uniform(-0.02, 0.02, (n, 12))
    s = pl.TrajectorySampler(ctl, pl.make_gait_plan("mini_cheetah", "walk", goal=(0.75, 0.0), base_height=bh))
    r = rollout(ctl, s, "id", q0, np.zeros((n, 18)), np.zeros(n), 80, 5e-3)
    assert (r.status_or == 0).all() and np.isfinite(r.q).all()
    s2 = pl.TrajectorySampler(ctl, pl.make_gait_plan("mini_cheetah", "trot", goal=(1.5, 0.0), base_height=bh))
    r = rollout(ctl, s2, "id", q0, np.zeros((n, 18)), np.zeros(n), 400, 5e-3)
    assert np.isfinite(r.q).all() and np.isfinite(r.v).all() and np.abs(r.v).max() < 1e6
    assert (r.status_or != 0).any()                              # straight-line base + 0.5 s diagonal supports: robots tip over
    assert np.allclose(r.t, 2.0)
