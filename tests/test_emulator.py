"""The exact device source (csrc/wbc_device.cuh) compiled for the host with the lock-step warp emulator
(tests/emu) against the oracle. Exercises the kernel's algorithm on CPU; the product path never uses it."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle import controllers as oc

GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def emu(built):
    lib = C.CDLL(str(Path(__file__).parent / "emu" / "libwbc_emu.so"))
    lib.emu_dynamics.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 8
    lib.emu_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]
    return lib


N_EMU = 24     # instances per case through the emulator (one OS thread per lane: ~0.25 s per instance)


def first(g, n=N_EMU):
    """The first n instances of a golden file as a dict."""
    return {k: (g[k][:n] if g[k].ndim and len(g[k]) >= n and k not in ("dof_order", "params") else g[k]) for k in g.files}


def run_step(emu, robot, kind, g, dof_order="depth_first", **params):
    from quadruped_drake_b200 import load_robot
    from quadruped_drake_b200.capi import KINDS, WbcIO, make_params, np_ptr
    ms, pr = load_robot(robot, dof_order=dof_order).as_struct(), make_params(**params)
    q, v, traj, contact = (np.ascontiguousarray(g[k]) for k in ("q", "v", "traj", "contact"))
    n = len(q)
    tau, met, st = np.zeros((n, 12)), np.zeros((n, 4)), np.zeros(n, np.int32)
    vd, f, qi = np.zeros((n, 18)), np.zeros((n, 4, 3)), np.zeros((n, 4))
    io = WbcIO(np_ptr(q), np_ptr(v), np_ptr(traj), np_ptr(contact), np_ptr(tau), np_ptr(met), np_ptr(st), np_ptr(vd), np_ptr(f), np_ptr(qi))
    assert emu.emu_step(C.byref(ms), C.byref(pr), KINDS[kind], n, C.byref(io)) == 0     # reduce -> record -> solve, as on the GPU
    return tau, met, st, vd, f, qi


@pytest.mark.parametrize("case", ["cfg2_mini_cheetah_stand", "cfg3_anymal_trot", "mixed_mini_cheetah"])
def test_emulated_dynamics_match_golden(emu, case):
    from quadruped_drake_b200 import load_robot
    from quadruped_drake_b200.capi import np_ptr
    g = first(np.load(GOLD / f"{case}.npz"))
    robot = "anymal_b" if "anymal" in case else "mini_cheetah"
    ms = load_robot(robot).as_struct()
    q, v = np.ascontiguousarray(g["q"]), np.ascontiguousarray(g["v"])
    n = len(q)
    M, Cv, tg = np.zeros((n, 18, 18)), np.zeros((n, 18)), np.zeros((n, 18))
    J, Jdv, pf = np.zeros((n, 4, 3, 18)), np.zeros((n, 4, 3)), np.zeros((n, 4, 3))
    emu.emu_dynamics(C.byref(ms), n, np_ptr(q), np_ptr(v), np_ptr(M), np_ptr(Cv), np_ptr(tg), np_ptr(J), np_ptr(Jdv), np_ptr(pf))
    for name, got in (("M", M), ("Cv", Cv), ("tau_g", tg), ("J_feet", J), ("Jdv_feet", Jdv), ("p_feet", pf)):
        ref = g[name][:n]
        m = len(ref)
        scale = np.abs(ref).reshape(m, -1).max(axis=1).reshape((m,) + (1,) * (ref.ndim - 1))
        assert (np.abs(got[:m] - ref) / np.maximum(scale, 1e-3)).max() < 1e-9, name     # 1e-9 relative (north star)


@pytest.mark.parametrize("case", ["cfg2_mini_cheetah_stand", "cfg3_anymal_trot", "cfg4_mini_cheetah_walk", "mixed_mini_cheetah"])
def test_emulated_id_step_matches_golden(emu, case):
    g = first(np.load(GOLD / f"{case}.npz"))
    robot = "anymal_b" if "anymal" in case else "mini_cheetah"
    tau, met, st, vd, f, qi = run_step(emu, robot, "id", g)
    assert (st == 0).all()
    assert np.abs(tau - g["id_tau"]).max() < 1e-5          # north star: 1e-5 on QP torques
    assert np.abs(vd - g["id_vd"]).max() < 1e-6
    assert np.abs(f - g["id_f"]).max() < 1e-5
    assert np.abs(qi[:, 0] - g["id_objective"]).max() < 1e-6 * max(1.0, np.abs(g["id_objective"]).max())
    assert np.abs(met[:, 1] - g["id_metrics"][:, 1]).max() < 1e-12


@pytest.mark.parametrize("kind", ["clf", "pc", "mptc"])
@pytest.mark.parametrize("case", ["cfg3_anymal_trot", "cfg4_mini_cheetah_walk"])
def test_emulated_clf_pc_steps_match_golden(emu, case, kind):
    g = first(np.load(GOLD / f"{case}.npz"), 12)
    robot = "anymal_b" if "anymal" in case else "mini_cheetah"
    tau, met, st, vd, f, qi = run_step(emu, robot, kind, g)
    ok = g[f"{kind}_ok"]
    assert (st[ok] == 0).all()
    if kind in ("pc", "mptc"):
        assert (st[~ok] == 64).all()                        # full flight: WBC_ST_UNSUPPORTED (the reference raises)
    assert np.abs(tau - g[f"{kind}_tau"])[ok].max() < 1e-5
    assert np.abs(vd - g[f"{kind}_vd"])[ok].max() < 1e-6
    assert np.abs(f - g[f"{kind}_f"])[ok].max() < 1e-5
    ref = g[f"{kind}_metrics"][ok]
    assert (np.abs(met[ok] - ref)[:, [0, 1, 3]] / np.maximum(1.0, np.abs(ref[:, [0, 1, 3]]))).max() < 1e-8


@pytest.mark.parametrize("case,kind", [("bf_mini_cheetah_mixed", "id"), ("bf_anymal_trot", "clf"), ("bf_mini_cheetah_mixed", "pc"),
                                       ("tl_mini_cheetah_walk", "id"), ("tl_anymal_trot", "id"), ("fixtures_mini_cheetah", "id"),
                                       ("fixtures_mini_cheetah", "clf")])
def test_emulated_other_configs_match_golden(emu, case, kind):
    """Breadth-first (2021-era Drake) velocity numbering, the torque box and the reference's manual test motions."""
    g = first(np.load(GOLD / f"{case}.npz"), 12)
    robot = "anymal_b" if "anymal" in case else "mini_cheetah"
    extra = {"torque_limits": 1} if case.startswith("tl_") else {}
    tau, met, st, vd, f, qi = run_step(emu, robot, kind, g, dof_order=str(g["dof_order"]), **extra)
    ok = g[f"{kind}_ok"]
    assert ok.sum() >= 9 and (st[ok] == 0).all()
    assert np.abs(tau - g[f"{kind}_tau"])[ok].max() < 1e-5
    assert np.abs(vd - g[f"{kind}_vd"])[ok].max() < 1e-6
    assert np.abs(f - g[f"{kind}_f"])[ok].max() < 1e-5
    assert (tau[st != 0] == 0).all()


def test_emulated_coriolis_matches_oracle(emu):
    from oracle.dynamics import Plant
    from quadruped_drake_b200 import load_robot
    from quadruped_drake_b200.capi import np_ptr
    emu.emu_coriolis.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4
    g = np.load(GOLD / "mixed_mini_cheetah.npz")
    P, ms = Plant("mini_cheetah"), load_robot("mini_cheetah").as_struct()
    n = 3
    q, v = np.ascontiguousarray(g["q"][:n]), np.ascontiguousarray(g["v"][:n])
    Cm, Jd = np.zeros((n, 18, 18)), np.zeros((n, 4, 3, 18))
    emu.emu_coriolis(C.byref(ms), n, np_ptr(q), np_ptr(v), np_ptr(Cm), np_ptr(Jd))
    for i in range(n):
        Co = P.coriolis_matrix(q[i], v[i])
        assert np.abs(Cm[i] - Co).max() < 1e-9 * np.abs(Co).max()
        assert np.abs(Cm[i] @ v[i] - g["Cv"][i]).max() < 1e-9 * max(1.0, np.abs(g["Cv"][i]).max())
        for k, fr in enumerate(P.foot_frames):
            assert np.abs(Jd[i, k] - P.frame_jacobian_dot(q[i], v[i], fr)).max() < 1e-9
