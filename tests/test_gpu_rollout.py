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


# ------------------------------------------------------------------------------------------ ground-contact plant
def grounded(ctl, n, rng=None, jitter=0.0, lift=0.0):
    """n copies of simulate.py's q0 with the feet exactly on the ground (+ optional joint jitter / lift)."""
    q0 = np.tile(Q0, (n, 1))
    if rng is not None and jitter > 0:
        q0[:, 7:] += rng.uniform(-jitter, jitter, (n, 12))
    pz = ctl.dynamics(q0, np.zeros((n, 18)))["p_feet"][:, :, 2]
    q0[:, 6] -= pz.min(axis=1)
    q0[:, 6] += lift
    return q0


def test_plant_step_matches_oracle(ctl):
    """wbc_plant_step (forward dynamics from tau + projected Gauss-Seidel ground contact + semi-implicit Euler) against the
    numpy restatement oracle/rollout.py:plant_step: standing, airborne, sliding, penetrating and random states."""
    from oracle import rollout as ro
    from oracle.dynamics import Plant
    from quadruped_drake_b200.rollout import plant_step
    P = Plant("mini_cheetah")
    rng = np.random.default_rng(7)
    n = 24
    q = grounded(ctl, n, rng, 0.2)
    v = rng.uniform(-1, 1, (n, 18)); v[:4] = 0.0
    tau = rng.uniform(-4, 4, (n, 12))
    q[1, 6] += 0.1; v[2, 3:5] = [0.8, -0.4]; q[3, 6] -= 0.004
    qn, vn, f, st = plant_step(ctl, q, v, tau, 5e-3)
    assert (st == 0).all()
    for i in range(n):
        qo, vo, fo = ro.plant_step(P, q[i], v[i], tau[i], 5e-3)
        assert np.abs(qn[i] - qo).max() < 1e-9 and np.abs(vn[i] - vo).max() < 1e-8, i
        assert np.abs(f[i] - fo).max() < 1e-6 * max(1.0, np.abs(fo).max())
    assert np.abs(f[1]).max() == 0.0                      # airborne robot: no ground force
    mu = 1.0
    assert (np.abs(f[:, :, 0]) <= mu * f[:, :, 2] + 1e-9).all() and (np.abs(f[:, :, 1]) <= mu * f[:, :, 2] + 1e-9).all() and (f[:, :, 2] >= 0).all()


def test_closed_loop_on_the_ground_stands_lands_and_lifts_a_foot(ctl):
    """The controller against the simulated robot on the ground (plant=True: torques in, contact from geometry). 1200 steps =
    the reference's 6 s: (a) a standing robot stays put and the ground carries its weight; (b) a robot dropped from 3 cm lands
    without going through the floor and settles on the reference; (c) RaiseFoot: the foot planned in swing really lifts and
    carries no force while the other three carry the weight."""
    import torch
    from quadruped_drake_b200 import planner as pl
    from quadruped_drake_b200.rollout import rollout
    rng = np.random.default_rng(11)
    n = 512
    q0 = grounded(ctl, n, rng, 0.03)
    bh = float(grounded(ctl, 1)[0, 6])
    q0[n // 2:, 6] += 0.03                                      # second half: dropped from 3 cm
    plans = [pl.make_motion_plan("mini_cheetah", "standing", 6.0, base_height=bh), pl.make_motion_plan("mini_cheetah", "raise_foot", 6.0, base_height=bh)]
    s = pl.TrajectorySampler(ctl, plans)
    pi = (np.arange(n) % 2).astype(np.int32)
    q, v, t, tpi = (torch.from_numpy(x).cuda() for x in (q0.copy(), np.zeros((n, 18)), np.zeros(n), pi))
    with torch.cuda.stream(torch.cuda.Stream()):
        r = rollout(ctl, s, "id", q, v, t, 1200, 5e-3, plan_index=tpi, plant=True)
    torch.cuda.synchronize()
    st = r.status_or.cpu().numpy()
    assert (st == 0).all(), np.unique(st, return_counts=True)
    qf, vf, f = q.cpu().numpy(), v.cpu().numpy(), r.f_contact.cpu().numpy()
    d = ctl.dynamics(qf, vf)
    pz = d["p_feet"][:, :, 2]
    mg = ctl.model.total_mass * 9.81
    assert pz.min() > -2e-3                                     # no penetration beyond the solver's tolerance
    assert np.abs(f[:, :, 2].sum(axis=1) - mg).max() < 0.02 * mg and np.abs(vf).max() < 0.05
    ref = s.sample(np.full(n, 6.0 - 1e-9), pi)["traj"]
    assert np.abs(qf[:, 4:7] - ref[:, 0:3]).max() < 0.01        # within 1 cm of the reference base position
    stand, lift = pi == 0, pi == 1
    assert pz[stand].max() < 2e-3 and (f[stand, :, 2] > 0.05 * mg).all()
    assert (pz[lift, 1] > 0.05).all() and np.abs(f[lift, 1]).max() == 0.0 and (f[lift][:, [0, 2, 3], 2] > 0.05 * mg).all()


def test_plant_rollout_shows_physical_failure_and_graph_equals_plain(ctl):
    """With the ground plant a dynamically inconsistent plan makes robots fall instead of being frozen at a failed QP: the
    tracking error grows although the QPs solve. Graph replay and plain launches give the same bits. The PD law of
    BasicController (basic_controller.py:322-352), which has no accelerations to integrate, runs against the plant too."""
    import torch
    from quadruped_drake_b200 import planner as pl
    from quadruped_drake_b200.rollout import rollout
    n = 256
    rng = np.random.default_rng(13)
    q0 = grounded(ctl, n, rng, 0.02)
    bh = float(grounded(ctl, 1)[0, 6])
    s = pl.TrajectorySampler(ctl, pl.make_gait_plan("mini_cheetah", "trot", goal=(1.5, 0.0), base_height=bh))
    outs = []
    for graph in (False, True):
        q, v, t = (torch.from_numpy(x).cuda() for x in (q0.copy(), np.zeros((n, 18)), np.zeros(n)))
        with torch.cuda.stream(torch.cuda.Stream()):
            r = rollout(ctl, s, "id", q, v, t, 300, 5e-3, use_graph=graph, plant=True)
        torch.cuda.synchronize()
        outs.append([x.cpu().numpy() for x in (q, v, r.tau, r.err_max, r.status_or, r.f_contact)])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    qf, vf, _, err_max, st, f = outs[0]
    assert np.isfinite(qf).all() and np.isfinite(vf).all()
    assert err_max.max() > 1e-3                                 # the straight-line trot plan is not trackable: errors grow
    assert (np.abs(f[:, :, 0]) <= f[:, :, 2] + 1e-9).all()
    # joint-space PD about the standing posture with the reference gains (Kp 30, Kd 1.5). The torque is held over the step
    # (explicit), so Kd dt / I_joint must stay below 2: with the 1e-3 kg m^2 knee links that needs dt = 1e-3, not 5e-3
    s2 = pl.TrajectorySampler(ctl, pl.make_motion_plan("mini_cheetah", "standing", 2.0, base_height=bh))
    r = rollout(ctl, s2, "pd", q0, np.zeros((n, 18)), np.zeros(n), 1000, 1e-3, plant=True)
    mg = ctl.model.total_mass * 9.81
    assert np.isfinite(r.q).all() and (r.status_or == 0).all() and r.q[:, 6].min() > 0.2 and np.abs(r.v).max() < 0.2
    assert np.abs(r.f_contact[:, :, 2].sum(axis=1) - mg).max() < 0.05 * mg        # the soft PD sags a little but stands
    r = rollout(ctl, s2, "pd", q0, np.zeros((n, 18)), np.zeros(n), 200, 5e-3, plant=True)
    assert np.isfinite(r.q).all() and np.isfinite(r.v).all()                       # beyond the limit: flagged and frozen, no NaNs
    assert (r.status_or[r.status_or != 0] == 128).all()


def test_reference_simulation_loop_with_the_dropin_pieces(built):
    """BASELINE configs[0]: the loop of the reference's simulate.py (planner -> IDController through its LeafSystem ports ->
    plant) for ONE robot, with the drop-in planner / controller mirrors and the ground-contact plant instead of Drake:
    2 s of standing and of the OrientationTest motion; the robot stands on the ground and follows the reference."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "examples"))
    import simulate
    q, v, log = simulate.run("id", "standing", sim_time=2.0, verbose=False)
    assert abs(q[6] - 0.3) < 5e-3 and np.abs(v).max() < 1e-2 and abs(log[-1, 3] - 8.252 * 9.81) < 1.0 and log[-1, 2] < 1e-3   # SimpleStanding: body at 0.3 m
    q, v, log = simulate.run("clf", "orientation", sim_time=2.0, verbose=False)
    t = 2.0 - 5e-3
    rpy = quat_to_rpy(q[None, :4])[0]
    assert np.abs(rpy - np.array([0.0, 0.4 * np.sin(t), 0.4 * np.cos(t)])).max() < 0.05 and log[-1, 2] < 5e-3
