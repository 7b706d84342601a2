"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle's golden vectors and,
at the benchmark's full sizes, through size-independent properties (KKT conditions of the reference QP)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
CASES = ["cfg2_mini_cheetah_stand", "cfg3_anymal_trot", "cfg4_mini_cheetah_walk", "mixed_mini_cheetah"]
# 2021-era Drake velocity numbering, the optional torque box with ACTIVE limits (BASELINE configs[2]) and the
# reference's manual test motions (planners/simple.py:87-115); every file holds >= 256 instances (tools/make_golden.py)
MORE = ["bf_mini_cheetah_mixed", "bf_anymal_trot", "tl_mini_cheetah_walk", "tl_anymal_trot", "fixtures_mini_cheetah"]


def robot_of(case):
    return "anymal_b" if "anymal" in case else "mini_cheetah"


def ctl_of(ctl_cache, case):
    """Controller handle built the way the golden file says (dof order, torque box)."""
    kw = {}
    if case.startswith("bf_"):
        kw["dof_order"] = "breadth_first"
    if case.startswith("tl_"):
        kw["torque_limits"] = 1
    return ctl_cache(robot_of(case), **kw)


@pytest.fixture(scope="module")
def ctl_cache(built):
    from quadruped_drake_b200.controller import BatchedController
    cache = {}

    def get(robot, dof_order="depth_first", **params):
        key = (robot, dof_order, tuple(sorted(params.items())))
        if key not in cache:
            cache[key] = BatchedController(robot, device=0, dof_order=dof_order, **params)
        return cache[key]
    yield get
    for c in cache.values():
        c.close()


@pytest.mark.parametrize("case", CASES + MORE)
def test_dynamics_match_golden(ctl_cache, case):
    """M, Cv, tau_g, J, Jdot v, p within 1e-9 relative (BASELINE.json north star). The dense terms (M, J) are stored for
    the first n_dyn instances of a file, the vectors for all of them."""
    g = np.load(GOLD / f"{case}.npz")
    d = ctl_of(ctl_cache, case).dynamics(g["q"], g["v"])
    for name in ("M", "Cv", "tau_g", "J_feet", "Jdv_feet", "p_feet"):
        ref = g[name]
        n = len(ref)
        got = d[name][:n]
        scale = np.abs(ref).reshape(n, -1).max(axis=1).reshape((n,) + (1,) * (ref.ndim - 1))
        assert (np.abs(got - ref) / np.maximum(scale, 1e-3)).max() < 1e-9, name


@pytest.mark.parametrize("case", CASES + MORE)
def test_id_step_matches_golden(ctl_cache, case):
    """tau within 1e-5 of the exact optimum of the reference QP (+ declared tie-break); vd, f, objective too.
    Instances the oracle could not certify (an infeasible torque box) are excluded and must not report success with
    non-zero torques either."""
    g = np.load(GOLD / f"{case}.npz")
    out = ctl_of(ctl_cache, case).step("id", g["q"], g["v"], g["traj"], g["contact"], debug=True)
    ok = g["id_ok"]
    assert ok.mean() > 0.95
    assert (out.status[ok] == 0).all(), np.unique(out.status[ok], return_counts=True)
    assert np.abs(out.tau - g["id_tau"])[ok].max() < 1e-5
    assert np.abs(out.vd - g["id_vd"])[ok].max() < 1e-6
    assert np.abs(out.f - g["id_f"])[ok].max() < 1e-5
    assert np.abs(out.qp_info[:, 0] - g["id_objective"])[ok].max() < 1e-6 * max(1.0, np.abs(g["id_objective"][ok]).max())
    assert np.abs(out.metrics[:, 1] - g["id_metrics"][:, 1])[ok].max() < 1e-12
    assert (out.tau[out.status != 0] == 0).all()
    if case.startswith("tl_"):
        lim = ctl_of(ctl_cache, case).model.effort[np.argsort(ctl_of(ctl_cache, case).model.act_index)]
        active = (np.abs(g["id_tau"]) > lim - 1e-7).any(axis=1) & ok
        assert active.sum() >= 25, "the torque-box goldens must exercise ACTIVE limits"
        assert (out.lam[active, 18:42] > 0).any(axis=1).all()


@pytest.mark.parametrize("case", CASES + MORE)
def test_clf_step_matches_golden(ctl_cache, case):
    """CLF-QP (clf_controller.py:48-234): the kernel's closed-form CARE constants against the oracle's numerical CARE."""
    g = np.load(GOLD / f"{case}.npz")
    out = ctl_of(ctl_cache, case).step("clf", g["q"], g["v"], g["traj"], g["contact"], debug=True)
    ok = g["clf_ok"]
    assert ok.mean() > 0.95
    assert (out.status[ok] == 0).all(), np.unique(out.status[ok], return_counts=True)
    assert np.abs(out.tau - g["clf_tau"])[ok].max() < 1e-5
    assert np.abs(out.vd - g["clf_vd"])[ok].max() < 1e-6
    assert np.abs(out.f - g["clf_f"])[ok].max() < 1e-5
    m, ref = out.metrics[ok], g["clf_metrics"][ok]
    scale = np.maximum(1.0, np.abs(ref))
    assert (np.abs(m[:, [0, 1, 3]] - ref[:, [0, 1, 3]]) / scale[:, [0, 1, 3]]).max() < 1e-8      # V, err, Vdot
    assert np.abs(out.qp_info[:, 0] - g["clf_objective"])[ok].max() < 1e-6 * max(1.0, np.abs(g["clf_objective"][ok]).max())


def test_clf_full_size_config4(ctl_cache):
    """BASELINE config 4 (CLF-QP, 65536 instances, walk contact patterns): all solve; reference constraints hold."""
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    q, v, traj, contact = generate(ctl.model, 65536, 20260121, "walk", ctl.fk)
    out = ctl.step("clf", q, v, traj, contact, debug=True)
    assert (out.status == 0).all(), np.unique(out.status, return_counts=True)
    dyn, con, fr = kkt_properties(ctl, q, v, traj, contact, out)
    assert dyn.max() < 1e-7 and con.max() < 1e-7 and fr.max() < 1e-7
    assert (out.qp_info[:, 2] > -1e-9).all()          # delta >= 0 is never optimal to violate: cost w*delta^2, row -delta
    assert_optimal("clf", ctl, q, v, traj, contact, out, sel=np.arange(0, 65536, 4))


@pytest.mark.parametrize("kind", ["pc", "mptc"])
@pytest.mark.parametrize("case", CASES + MORE)
def test_pc_step_matches_golden(ctl_cache, case, kind):
    """Passivity-constrained QP (pc_controller.py:43-255) and MPTC (mptc_controller.py:125-310): analytic C w / Jdot
    against the oracle's AutoDiff-equivalent."""
    g = np.load(GOLD / f"{case}.npz")
    if f"{kind}_tau" not in g.files:
        pytest.skip("golden file holds no MPTC vectors")
    out = ctl_of(ctl_cache, case).step(kind, g["q"], g["v"], g["traj"], g["contact"], debug=True)
    g = {k.replace(kind + "_", "pc_") if k.startswith(kind + "_") else k: g[k] for k in g.files if kind == "pc" or not k.startswith("pc_")}
    ok = g["pc_ok"]
    flight = g["contact"].sum(axis=1) == 0
    assert (out.status[ok] == 0).all() and (out.status[flight] == 64).all() and (ok | flight).mean() > 0.95
    assert np.abs(out.tau - g["pc_tau"])[ok].max() < 1e-5
    assert np.abs(out.vd - g["pc_vd"])[ok].max() < 1e-6
    assert np.abs(out.f - g["pc_f"])[ok].max() < 1e-5
    ref = g["pc_metrics"][ok]
    assert (np.abs(out.metrics[ok] - ref)[:, [0, 1, 3]] / np.maximum(1.0, np.abs(ref[:, [0, 1, 3]]))).max() < 1e-8


def test_pc_full_size_config4(ctl_cache):
    """BASELINE config 4 (PC variant, 65536 instances, walk patterns): solves; constraints and passivity row hold; a random
    sample of the batch (PC and MPTC) is the optimum of the oracle's full-size QP."""
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    q, v, traj, contact = generate(ctl.model, 65536, 20260121, "walk", ctl.fk)
    out = ctl.step("pc", q, v, traj, contact, debug=True)
    assert (out.status == 0).all(), np.unique(out.status, return_counts=True)
    dyn, con, fr = kkt_properties(ctl, q, v, traj, contact, out)
    assert dyn.max() < 1e-7 and con.max() < 1e-7 and fr.max() < 1e-7
    assert (out.metrics[:, 3] < 1e-7 * np.maximum(1.0, np.abs(out.metrics[:, 0]))).all()      # Vdot <= delta <= 0
    # optimality on a random sample of the full-size batch: the oracle's PC-QP (AutoDiff-equivalent C and Jdot, full-size QP)
    from oracle import controllers as oc
    mptc = ctl.step("mptc", q, v, traj, contact, debug=True)
    assert (mptc.status == 0).all()
    for got, octl, cnt in ((out, oc.PCController("mini_cheetah"), 24), (mptc, oc.MPTCController("mini_cheetah"), 12)):
        for i in np.random.default_rng(5).choice(65536, cnt, replace=False):
            o = octl.control_law(q[i], v[i], oc.traj_to_dict(traj[i], contact[i]))
            assert o.status in ("optimal", "ipm")
            assert np.abs(got.tau[i] - o.tau).max() < 1e-5 and np.abs(got.vd[i] - o.vd).max() < 1e-6, i


def test_coriolis_entry_matches_oracle(ctl_cache):
    """wbc_coriolis: C = 1/2 d(Cv)/dv (CalcCoriolisMatrix) and the four foot Jdot (CalcFrameJacobianDot), 1e-9 relative."""
    from oracle.dynamics import Plant
    for robot, case in (("mini_cheetah", "mixed_mini_cheetah"), ("anymal_b", "cfg3_anymal_trot")):
        g = np.load(GOLD / f"{case}.npz")
        ctl, P = ctl_cache(robot), Plant(robot)
        Cm, Jd = ctl.coriolis(g["q"], g["v"])
        assert (np.abs(np.einsum("nij,nj->ni", Cm, g["v"]) - g["Cv"]).max(axis=1) < 1e-9 * np.maximum(1, np.abs(g["Cv"]).max(axis=1))).all()
        for i in range(3):
            Co = P.coriolis_matrix(g["q"][i], g["v"][i])
            assert np.abs(Cm[i] - Co).max() < 1e-9 * np.abs(Co).max()
            for k, fr in enumerate(P.foot_frames):
                assert np.abs(Jd[i, k] - P.frame_jacobian_dot(g["q"][i], g["v"][i], fr)).max() < 1e-9


def test_pd_law(ctl_cache):
    """BasicController.ControlLaw (basic_controller.py:322-352) and its LeafSystem mirror."""
    from oracle import controllers as oc
    from quadruped_drake_b200.controller import BasicController
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    q, v, traj, contact = generate(ctl.model, 257, 9, "stand", ctl.fk)
    q[0, 7:] += 10.0                                            # drive the clip
    tau = ctl.step_pd(q, v)
    ref = oc.BasicController("mini_cheetah")
    for i in (0, 1, 100, 256):
        assert np.abs(tau[i] - ref.control_law(q[i], v[i])).max() < 1e-12
    assert np.abs(tau[0]).max() == 150.0
    leaf = BasicController("mini_cheetah", 5e-3)
    ctx = leaf.CreateDefaultContext()
    ctx.FixValue(0, np.hstack([q[1], v[1]]))
    assert np.abs(leaf.EvalOutput(ctx, 0) - tau[1]).max() == 0.0


def test_named_wrapper_and_torch_device_path(ctl_cache):
    """wbc_step_id on device pointers (torch only carries the memory and the stream) equals the host entry."""
    import ctypes as C
    import torch
    g = np.load(GOLD / "mixed_mini_cheetah.npz")
    ctl = ctl_cache("mini_cheetah")
    host = ctl.step("id", g["q"], g["v"], g["traj"], g["contact"])
    dev = torch.device("cuda:0")
    tq, tv, tt = (torch.from_numpy(np.ascontiguousarray(g[k])).to(dev) for k in ("q", "v", "traj"))
    tc = torch.from_numpy(np.ascontiguousarray(g["contact"])).to(dev)
    out = ctl.step("id", tq, tv, tt, tc)
    torch.cuda.synchronize()
    assert np.array_equal(out.tau.cpu().numpy(), host.tau)            # bit-identical: same kernel, same inputs
    n = len(g["q"])
    tau = torch.empty((n, 12), dtype=torch.float64, device=dev)
    met = torch.empty((n, 4), dtype=torch.float64, device=dev)
    st = torch.empty((n,), dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    rc = ctl.lib.wbc_step_id(ctl._h, n, p(tq), p(tv), p(tt), p(tc), p(tau), p(met), p(st), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0 and np.array_equal(tau.cpu().numpy(), host.tau)


def kkt_properties(ctl, q, v, traj, contact, out, mu=0.7, kd=100.0):  # noqa: E302
    """Residuals of the *reference* constraints evaluated with the GPU's own dynamics."""
    d = ctl.dynamics(q, v)
    B = ctl.model.actuation_matrix()
    n = len(q)
    c = contact.astype(bool)
    f = out.f * c[:, :, None]
    lhs = np.einsum("nij,nj->ni", d["M"], out.vd) + d["Cv"] + d["tau_g"]
    rhs = out.tau @ B.T + np.einsum("nkij,nki->nj", d["J_feet"], f)
    dyn = np.abs(lhs - rhs).max(axis=1)
    acc = np.einsum("nkij,nj->nki", d["J_feet"], out.vd) + d["Jdv_feet"] + kd * np.einsum("nkij,nj->nki", d["J_feet"], v)
    con = np.abs(acc * c[:, :, None]).reshape(n, -1).max(axis=1)
    fr = np.maximum(np.abs(f[:, :, 0]) - mu * f[:, :, 2], np.abs(f[:, :, 1]) - mu * f[:, :, 2]).max(axis=1)
    return dyn, con, fr


@pytest.mark.parametrize("robot,pattern,n,seed", [("mini_cheetah", "stand", 4096, 20260119), ("anymal_b", "trot", 16384, 20260120),
                                                   ("mini_cheetah", "mixed", 8192, 5)])
def test_full_size_properties(ctl_cache, robot, pattern, n, seed):
    """BASELINE configs 2 and 3 at full size: every instance solves, and the solution satisfies the reference's
    dynamics / contact / friction constraints to 1e-7 (absolute, on O(1..100) quantities)."""
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache(robot)
    q, v, traj, contact = generate(ctl.model, n, seed, pattern, ctl.fk)
    out = ctl.step("id", q, v, traj, contact, debug=True)
    assert (out.status == 0).all(), np.unique(out.status, return_counts=True)
    dyn, con, fr = kkt_properties(ctl, q, v, traj, contact, out)
    assert dyn.max() < 1e-7 and con.max() < 1e-7 and fr.max() < 1e-7
    # determinism / idempotence: same inputs, same bits
    again = ctl.step("id", q, v, traj, contact)
    assert np.array_equal(again.tau, out.tau)
    # swing feet carry no force, flight instances have tau = M_j vd + h_j only
    assert np.all(out.f[contact == 0] == 0)
    assert_optimal("id", ctl, q, v, traj, contact, out)


def assert_optimal(kind, ctl, q, v, traj, contact, out, params=None, sel=None):
    """Optimality, not just feasibility: with the exported multipliers the returned [vd; tau; f] satisfies the KKT conditions
    of the reference QP (tests/kkt.py) - a feasible but sub-optimal torque fails the stationarity row."""
    import kkt
    from types import SimpleNamespace
    if sel is not None:
        q, v, traj, contact = q[sel], v[sel], traj[sel], contact[sel]
        out = SimpleNamespace(tau=out.tau[sel], vd=out.vd[sel], f=out.f[sel], qp_info=out.qp_info[sel], lam=out.lam[sel])
    cert = kkt.certificate(kind, ctl.dynamics(q, v), ctl.model, q, v, traj, contact, out, params)
    assert cert["stationarity"].max() < 1e-6, cert["stationarity"].max()      # relative to the cost gradient
    assert cert["dual"].min() >= 0.0
    assert cert["comp"].max() < 1e-6, cert["comp"].max()
    assert cert["eq"].max() < 1e-7 and cert["ineq"].max() < 1e-7
    return cert


def test_full_size_config3_torque_limits_optimal(ctl_cache):
    """BASELINE configs[2] as specified: anymal_b trot, 16384 instances, friction pyramid + torque limits ON. Every instance
    either solves to a KKT point of the reference QP with the box rows, or reports a status and returns zero torques."""
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("anymal_b", torque_limits=1)
    q, v, traj, contact = generate(ctl.model, 16384, 20260120, "trot", ctl.fk)
    out = ctl.step("id", q, v, traj, contact, debug=True)
    ok = out.status == 0
    assert ok.mean() > 0.999, np.unique(out.status, return_counts=True)
    assert (out.tau[~ok] == 0).all() and (out.lam[~ok] == 0).all()
    lim = ctl.model.effort[np.argsort(ctl.model.act_index)]
    assert (np.abs(out.tau) <= lim + 1e-7).all()
    assert_optimal("id", ctl, q, v, traj, contact, out, {"torque_limits": 1}, sel=ok)


def test_failed_instances_return_zero_torques(ctl_cache):
    """include/wbc.h: any status bit => tau = f = vd = lam = 0 (the reference asserts instead). Failures provoked here: an
    iteration cap of 2 (MAXITER), a zero quaternion (BADQUAT), gimbal lock (GIMBAL), PC in full flight (UNSUPPORTED)."""
    from quadruped_drake_b200 import capi
    from quadruped_drake_b200.synth import generate
    capped = ctl_cache("mini_cheetah", max_iter=2)
    q, v, traj, contact = generate(capped.model, 512, 31, "stand", capped.fk)
    q[1, 0:4] = 0.0
    q[2, 0:4] = [np.cos(np.pi / 4), 0.0, np.sin(np.pi / 4), 0.0]
    out = capped.step("id", q, v, traj, contact, debug=True)
    bad = out.status != 0
    assert (out.status & capi.ST_MAXITER).astype(bool).sum() > 100 and out.status[1] & capi.ST_BADQUAT and out.status[2] & capi.ST_GIMBAL
    for arr in (out.tau, out.f, out.vd, out.lam):
        assert (arr[bad] == 0).all()
    good = ~bad
    ref = ctl_cache("mini_cheetah").step("id", q, v, traj, contact)
    assert good.any() and np.array_equal(out.tau[good], ref.tau[good])          # the cap does not touch instances below it
    contact[:8] = 0
    pc = ctl_cache("mini_cheetah").step("pc", q, v, traj, contact, debug=True)
    assert (pc.status[3:8] == capi.ST_UNSUPPORTED).all() and (pc.tau[:8] == 0).all()


def test_instance_independence_and_ragged_sizes(ctl_cache):
    """Results do not depend on batch size or position in the batch (N = 1, 3, 5, 127 incl. partial CTAs)."""
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    q, v, traj, contact = generate(ctl.model, 127, 77, "mixed", ctl.fk)
    full = ctl.step("id", q, v, traj, contact).tau
    for n in (1, 3, 5):
        assert np.array_equal(ctl.step("id", q[:n], v[:n], traj[:n], contact[:n]).tau, full[:n])
    perm = np.random.default_rng(0).permutation(127)
    assert np.array_equal(ctl.step("id", q[perm], v[perm], traj[perm], contact[perm]).tau, full[perm])
    empty = ctl.step("id", q[:0], v[:0], traj[:0], contact[:0])
    assert empty.tau.shape == (0, 12)


def test_status_flags(ctl_cache):
    """Per-instance failure reporting instead of the reference's assert: bad quaternion, gimbal lock."""
    from quadruped_drake_b200 import capi
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    q, v, traj, contact = generate(ctl.model, 8, 3, "stand", ctl.fk)
    q[1, 0:4] = 0.0                                             # zero quaternion
    q[2, 0:4] = [np.cos(np.pi / 4), 0.0, np.sin(np.pi / 4), 0.0]  # pitch = +90 deg
    out = ctl.step("id", q, v, traj, contact)
    assert out.status[1] & capi.ST_BADQUAT
    assert out.status[2] & capi.ST_GIMBAL
    assert (out.status[[0, 3, 4, 5, 6, 7]] == 0).all()
    assert np.isfinite(out.tau[[0, 3, 4, 5, 6, 7]]).all()


def test_torque_limits_option(ctl_cache):
    """Optional |tau| <= URDF effort box (not in the reference QP): the box holds, an inactive box changes nothing,
    an active one is a KKT point of the QP with the box rows, and whatever does not solve is flagged and zeroed."""
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah", torque_limits=1)
    free = ctl_cache("mini_cheetah")
    q, v, traj, contact = generate(ctl.model, 2048, 21, "walk", ctl.fk)
    lim = ctl.model.effort[np.argsort(ctl.model.act_index)]
    a, b = ctl.step("id", q, v, traj, contact, debug=True), free.step("id", q, v, traj, contact)
    okay = a.status == 0
    assert okay.mean() > 0.995, np.unique(a.status, return_counts=True)
    assert (a.tau[~okay] == 0).all()
    assert (np.abs(a.tau[okay]) <= lim + 1e-7).all()
    inside = okay & (np.abs(b.tau) < lim - 1e-3).all(axis=1)
    assert inside.any() and np.abs(a.tau[inside] - b.tau[inside]).max() < 1e-6   # inactive box changes nothing
    active = okay & (a.lam[:, 18:42] > 0).any(axis=1)
    assert active.sum() > 100
    assert_optimal("id", ctl, q, v, traj, contact, a, {"torque_limits": 1}, sel=okay)


def test_leafsystem_mirror_standing(built):
    """Config 1: the reference port layout for one instance (simulate.py:106-139 wiring, SimpleStanding input)."""
    from oracle import controllers as oc
    from quadruped_drake_b200.controller import IDController
    ctl = IDController("mini_cheetah", 5e-3)
    assert [ctl.get_input_port(i).get_name() for i in (0, 1)] == ["quad_state", "trunk_input"]
    assert [ctl.get_output_port(i).get_name() for i in (0, 1)] == ["quad_torques", "output_metrics"]
    q0 = np.array([1, 0, 0, 0, 0, 0, 0.3] + [0, -0.8, 1.6] * 4, float)
    ctx = ctl.CreateDefaultContext()
    ctx.FixValue(0, np.hstack([q0, np.zeros(18)]))
    ctx.FixValue(1, oc.standing_dict())
    tau = ctl.EvalOutput(ctx, 0)
    ref = oc.IDController("mini_cheetah").control_law(q0, np.zeros(18), oc.standing_dict())
    assert np.abs(tau - ref.tau).max() < 1e-5
    met = ctl.EvalOutput(ctx, 1)
    assert abs(met[1] - ref.metrics[1]) < 1e-12


@pytest.mark.gpu
def test_host_entry_paths_agree(ctl_cache):
    """wbc_step_host: pageable buffers (staged, two-stream chunked copies) and page-locked buffers (zero-copy: the kernel
    reads / writes host memory directly; from 24576 instances on the inputs go through the copy engine in four chunks and only
    the outputs are written zero-copy) give bit-identical results, for sizes around the chunking thresholds."""
    import ctypes as C
    from quadruped_drake_b200 import capi
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    for n in (1, 5, 2047, 2048, 4099, 24575, 24581):
        q, v, traj, contact = generate(ctl.model, n, 77 + n, "mixed", ctl.fk)
        a = ctl.step("id", q, v, traj, contact, debug=True)                      # numpy arrays: pageable path
        hq, hv, ht = capi.pinned_empty((n, 19)), capi.pinned_empty((n, 18)), capi.pinned_empty((n, 54))
        hc = capi.pinned_empty((n, 4), np.uint8)
        hq[:], hv[:], ht[:], hc[:] = q, v, traj, contact
        tau, met, st = capi.pinned_empty((n, 12)), capi.pinned_empty((n, 4)), capi.pinned_empty((n,), np.int32)
        vd = capi.pinned_empty((n, 18))
        io = capi.WbcIO(capi.np_ptr(hq), capi.np_ptr(hv), capi.np_ptr(ht), capi.np_ptr(hc), capi.np_ptr(tau), capi.np_ptr(met),
                        capi.np_ptr(st), capi.np_ptr(vd), None, None)
        assert ctl.lib.wbc_step_host(ctl._h, capi.WBC_CTRL_ID, n, C.byref(io)) == 0
        assert np.array_equal(tau, a.tau) and np.array_equal(met, a.metrics) and np.array_equal(st, a.status) and np.array_equal(vd, a.vd)


@pytest.mark.gpu
def test_large_batch_chunks_and_staged_pipeline(ctl_cache):
    """More than one reduce / solve launch pair (262144 instances per pair) and, on page-locked host buffers, the staged
    two-stream copy pipeline that large batches take: bit-identical to the same instances solved in a small batch."""
    import ctypes as C
    from quadruped_drake_b200 import capi
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    base = 4096
    q0, v0, t0, c0 = generate(ctl.model, base, 91, "mixed", ctl.fk)
    small = ctl.step("id", q0, v0, t0, c0)
    n = 262144 + 4096 + 5                                         # two launch pairs, ragged tail
    idx = np.arange(n) % base
    hq, hv, ht = capi.pinned_empty((n, 19)), capi.pinned_empty((n, 18)), capi.pinned_empty((n, 54))
    hc = capi.pinned_empty((n, 4), np.uint8)
    hq[:], hv[:], ht[:], hc[:] = q0[idx], v0[idx], t0[idx], c0[idx]
    tau, met, st = capi.pinned_empty((n, 12)), capi.pinned_empty((n, 4)), capi.pinned_empty((n,), np.int32)
    io = capi.WbcIO(capi.np_ptr(hq), capi.np_ptr(hv), capi.np_ptr(ht), capi.np_ptr(hc), capi.np_ptr(tau), capi.np_ptr(met),
                    capi.np_ptr(st), None, None, None)
    launches0 = ctl.launches
    assert ctl.lib.wbc_step_host(ctl._h, capi.WBC_CTRL_ID, n, C.byref(io)) == 0
    assert ctl.launches - launches0 >= 4
    assert np.array_equal(tau, small.tau[idx]) and np.array_equal(st, small.status[idx]) and np.array_equal(met, small.metrics[idx])


@pytest.mark.gpu
def test_unaligned_buffers_and_kernel_profile(ctl_cache):
    """Rows that do not start on a 16-byte boundary take the per-lane input path of the reduce kernel (no bulk staging):
    same bits. wbc_profile_step splits a step into its two kernels."""
    import ctypes as C
    import torch
    from quadruped_drake_b200 import capi
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    n = 1024
    q, v, traj, contact = generate(ctl.model, n, 17, "walk", ctl.fk)
    ref = ctl.step("id", q, v, traj, contact)
    dev = torch.device("cuda:0")

    def shifted(a, dtype):                       # device copy whose first element sits 8 bytes past a 16-byte boundary
        flat = torch.empty(a.size + 16, dtype=dtype, device=dev)
        off = 1 if dtype == torch.float64 else 8
        view = flat[off:off + a.size]
        view.copy_(torch.from_numpy(np.ascontiguousarray(a).reshape(-1)))
        assert view.data_ptr() % 16 == 8
        return view
    tq, tv, tt = shifted(q, torch.float64), shifted(v, torch.float64), shifted(traj, torch.float64)
    tc = shifted(contact, torch.uint8)
    tau = torch.empty((n, 12), dtype=torch.float64, device=dev)
    met = torch.empty((n, 4), dtype=torch.float64, device=dev)
    st = torch.empty((n,), dtype=torch.int32, device=dev)
    io = capi.WbcIO(tq.data_ptr(), tv.data_ptr(), tt.data_ptr(), tc.data_ptr(), tau.data_ptr(), met.data_ptr(), st.data_ptr(), None, None, None)
    stream = torch.cuda.current_stream(dev)
    assert ctl.lib.wbc_step(ctl._h, capi.WBC_CTRL_ID, n, C.byref(io), C.c_void_p(stream.cuda_stream)) == 0
    torch.cuda.synchronize()
    assert np.array_equal(tau.cpu().numpy(), ref.tau) and np.array_equal(st.cpu().numpy(), ref.status)
    ms_r, ms_s = C.c_double(), C.c_double()
    assert ctl.lib.wbc_profile_step(ctl._h, capi.WBC_CTRL_ID, n, C.byref(io), 5, C.c_void_p(stream.cuda_stream), C.byref(ms_r), C.byref(ms_s)) == 0
    assert 0.0 < ms_r.value < 10.0 and 0.0 < ms_s.value < 10.0


def test_multi_gpu_call_matches_single_handle(ctl_cache):
    """wbc_multi_step_host (MultiGpuController): one call, contiguous shards on every listed device, host arrays hold all
    results. On a one-GPU box the two shards run on two handles of the same device; with >= 2 devices on all of them."""
    import torch
    from quadruped_drake_b200.sharding import MultiGpuController, shard_range
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    n = 10007                                                    # ragged: shards of different sizes
    q, v, traj, contact = generate(ctl.model, n, 41, "mixed", ctl.fk)
    ref = ctl.step("id", q, v, traj, contact)
    ndev = torch.cuda.device_count()
    for devices in ([0, 0], list(range(ndev)), [0] * 3 + list(range(ndev))):
        multi = MultiGpuController("mini_cheetah", devices=devices)
        buf = multi.pinned(n)
        buf["q"][:], buf["v"][:], buf["traj"][:], buf["contact"][:] = q, v, traj, contact
        l0 = multi.launches
        out = multi.step("id", buf["q"], buf["v"], buf["traj"], buf["contact"], buf["tau"], buf["metrics"], buf["status"])
        assert np.array_equal(out.tau, ref.tau) and np.array_equal(out.status, ref.status) and np.array_equal(out.metrics, ref.metrics)
        assert multi.launches - l0 >= 2 * len(devices)
        pag = multi.step("clf", q, v, traj, contact)             # pageable arrays: staged per shard
        assert np.array_equal(pag.tau, ctl.step("clf", q, v, traj, contact).tau)
        assert shard_range(n, len(devices) - 1, len(devices))[1] == n
        small = multi.step("id", q[:2], v[:2], traj[:2], contact[:2])      # fewer instances than devices: empty shards
        assert np.array_equal(small.tau, ref.tau[:2])
        multi.close()


def test_leafsystem_batched_ports(built):
    """for N > 1 the same port names carry instance-major vectors of width 37N / 12N / 4N and the abstract port a
    dict of arrays with a leading N axis; a failing instance raises like the reference's assert."""
    from oracle import controllers as oc
    from quadruped_drake_b200.controller import BasicController, CLFController, IDController
    n = 5
    ctl = IDController("mini_cheetah", 5e-3, n_instances=n)
    assert ctl.get_input_port(0).model_value.size() == 37 * n and ctl.get_output_port(0).model_value.size() == 12 * n
    assert ctl.get_output_port(1).model_value.size() == 4 * n
    g = np.load(GOLD / "fixtures_mini_cheetah.npz")
    q, v, traj, contact = g["q"][:n], g["v"][:n], g["traj"][:n], g["contact"][:n]
    dicts = [oc.traj_to_dict(traj[i], contact[i]) for i in range(n)]
    batch = {k: np.stack([np.asarray(d[k]) for d in dicts]) for k in dicts[0] if k not in ("f_cj", "u2_max")}
    ctx = ctl.CreateDefaultContext()
    ctx.FixValue(0, np.hstack([q, v]).ravel())
    ctx.FixValue(1, batch)
    tau = ctl.EvalOutput(ctx, 0).reshape(n, 12)
    assert np.abs(tau - g["id_tau"][:n]).max() < 1e-5
    met = ctl.EvalOutput(ctx, 1).reshape(n, 4)
    assert np.abs(met[:, 1] - g["id_metrics"][:n, 1]).max() < 1e-12
    one = CLFController("mini_cheetah", 5e-3)                    # N = 1 keeps the reference's scalar logging attributes
    c1 = one.CreateDefaultContext()
    c1.FixValue(0, np.hstack([q[0], v[0]]))
    c1.FixValue(1, dicts[0])
    assert np.abs(one.EvalOutput(c1, 0) - g["clf_tau"][0]).max() < 1e-5 and isinstance(one.V, float)
    bad = np.hstack([q, v])
    bad[2, 0:4] = 0.0                                            # zero quaternion on instance 2
    ctx.FixValue(0, bad.ravel())
    with pytest.raises(AssertionError, match="instance"):
        ctl.EvalOutput(ctx, 0)
    pd = BasicController("mini_cheetah", 5e-3)
    assert [p.get_name() for p in pd._in] == ["quad_state"]      # no trunk input on the PD controller (basic_controller.py:33-50)


def test_towr_planner_u2_max(ctl_cache):
    """TowrTrunkPlanner.ComputeMaxControlInputs (planners/towr.py:70-90): max |[foot pdd; rpydd; pdd]| over the stored samples,
    attached to the dict once the motion has started."""
    from quadruped_drake_b200 import planner as pl
    ctl = ctl_cache("mini_cheetah")
    p = pl.TowrTrunkPlanner(ctl, robot="mini_cheetah")
    o = p.sampler.sample(np.asarray(p.plan.grid) + p.wait_time)
    tr = o["traj"]
    ref = np.linalg.norm(np.concatenate([tr[:, 42:54], tr[:, 15:18], tr[:, 6:9]], axis=1), axis=1).max()
    assert p.u2_max == pytest.approx(ref) and p.u2_max > 0.0
    assert p.SetTrunkOutputs(0.5)["u2_max"] == 0.0 and p.SetTrunkOutputs(2.0)["u2_max"] == pytest.approx(ref)


def test_chunked_device_step_matches_single_chain(ctl_cache):
    """include/wbc.h: a 4096-65536 instance wbc_step is issued as 2-4 chunks on as many streams, joined to the caller's stream.
    Bit-identical to the same instances solved in pieces below the threshold (one reduce -> solve chain), ragged tails included,
    with the optional outputs (vd, f, lam) on, and complete when the caller's stream is."""
    import torch
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    dev = torch.device("cuda:0")
    for n in (4099, 6145, 8195):
        q, v, traj, contact = generate(ctl.model, n, 300 + n, "mixed", ctl.fk)
        pieces = [ctl.step("id", q[o:o + 2000], v[o:o + 2000], traj[o:o + 2000], contact[o:o + 2000], debug=True) for o in range(0, n, 2000)]
        t = [torch.from_numpy(x).to(dev) for x in (q, v, traj, contact)]
        st = torch.cuda.Stream()
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            out = ctl.step("id", *t, debug=True)
        st.synchronize()                                            # only the caller's stream: the internal ones must have joined it
        for name in ("tau", "metrics", "status", "vd", "f", "lam"):
            a = getattr(out, name).cpu().numpy().reshape(n, -1)
            b = np.concatenate([np.asarray(getattr(p_, name)).reshape(len(p_.status), -1) for p_ in pieces])
            assert np.array_equal(a, b), (n, name)


def test_steps_on_two_streams_of_one_handle_are_ordered(ctl_cache):
    """include/wbc.h: the hand-over scratch of a handle belongs to one step at a time. A step issued on another stream than
    the previous one is ordered behind it by the library (event dependency) instead of corrupting the scratch."""
    import ctypes as C
    import torch
    from quadruped_drake_b200 import capi
    from quadruped_drake_b200.synth import generate
    ctl = ctl_cache("mini_cheetah")
    n = 32768
    dev = torch.device("cuda:0")
    sets = []
    for seed in (1, 2):
        q, v, traj, contact = generate(ctl.model, n, seed, "stand", ctl.fk)
        ref = ctl.step("id", q, v, traj, contact).tau
        t = [torch.from_numpy(x).to(dev) for x in (q, v, traj)] + [torch.from_numpy(contact).to(dev)]
        out = [torch.empty((n, 12), dtype=torch.float64, device=dev), torch.empty((n, 4), dtype=torch.float64, device=dev),
               torch.empty((n,), dtype=torch.int32, device=dev)]
        sets.append((ctl.make_io(*t, *out), out, ref, t))
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(5):
        for (io, out, ref, _), st in zip(sets, (sa, sb)):
            assert ctl.lib.wbc_step(ctl._h, capi.WBC_CTRL_ID, n, C.byref(io), C.c_void_p(st.cuda_stream)) == 0
    torch.cuda.synchronize()
    for io, out, ref, _ in sets:
        assert np.array_equal(out[0].cpu().numpy(), ref)


def test_pc_infeasible_states_agree_with_the_oracle(ctl_cache):
    """The passivity row (Vdot <= delta, delta <= 0) can make the PC-QP genuinely infeasible (0.14 % of random anymal_b stand
    states in the soak run, profiles/r2_soak.jsonl); the reference asserts there. For 24 such states the oracle's exact solver finds
    no feasible point either (primal residual 5-235), the kernel reports WBC_ST_INFEASIBLE with zero torques; the 8 solvable
    neighbours agree to 1e-5."""
    from quadruped_drake_b200 import capi
    g = np.load(GOLD / "pc_infeasible_anymal.npz")
    out = ctl_cache("anymal_b").step("pc", g["q"], g["v"], g["traj"], g["contact"])
    ok = g["pc_ok"]
    assert (~ok).sum() == 24 and (g["oracle_primal_res"][~ok] > 1.0).all()
    assert (out.status[~ok] == capi.ST_INFEASIBLE).all() and (out.tau[~ok] == 0).all()
    assert (out.status[ok] == 0).all() and np.abs(out.tau - g["pc_tau"])[ok].max() < 1e-5
