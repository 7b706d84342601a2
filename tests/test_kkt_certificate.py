"""The batched KKT certificate (tests/kkt.py) against the oracle on CPU: it accepts the oracle's optimum with the oracle's
multipliers, rejects a feasible but sub-optimal point, and accepts what the device code (host warp emulator) exports."""
import ctypes as C
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

import kkt
from oracle import controllers as oc
from oracle.dynamics import Plant

GOLD = Path(__file__).parent / "golden"


def oracle_dynamics(plant, q, v):
    n = len(q)
    d = dict(M=np.zeros((n, 18, 18)), Cv=np.zeros((n, 18)), tau_g=np.zeros((n, 18)), J_feet=np.zeros((n, 4, 3, 18)),
             Jdv_feet=np.zeros((n, 4, 3)), p_feet=np.zeros((n, 4, 3)))
    for i in range(n):
        d["M"][i], d["Cv"][i], d["tau_g"][i], _ = plant.calc_dynamics(q[i], v[i])
        for k, f in enumerate(plant.foot_frames):
            d["p_feet"][i, k], d["J_feet"][i, k], d["Jdv_feet"][i, k] = plant.frame_position_quantities(q[i], v[i], f)
    return d


def oracle_outputs(kind, plant, q, v, traj, contact, **params):
    """Oracle step for every instance, multipliers rearranged into the lam layout of include/wbc.h."""
    ctl = {"id": oc.IDController, "clf": oc.CLFController}[kind](plant, **params)
    n = len(q)
    out = SimpleNamespace(tau=np.zeros((n, 12)), vd=np.zeros((n, 18)), f=np.zeros((n, 4, 3)), qp_info=np.zeros((n, 4)),
                          lam=np.zeros((n, 42)))
    for i in range(n):
        o = ctl.control_law(q[i], v[i], oc.traj_to_dict(traj[i], contact[i]))
        assert o.status == "optimal"
        out.tau[i], out.vd[i], out.f[i] = o.tau, o.vd, o.f
        lam, r = np.asarray(o.lam), 0
        if kind == "clf":
            out.lam[i, 16], r = lam[0], 1
            out.qp_info[i, 2] = o.delta
        for k in range(4):
            if contact[i, k]:
                out.lam[i, 4 * k:4 * k + 4] = lam[r:r + 4]
                r += 4
        if params.get("torque_limits"):
            out.lam[i, 18:42] = lam[r:r + 24]
    return out


@pytest.mark.parametrize("kind,case,params", [("id", "mixed_mini_cheetah", {}), ("clf", "cfg4_mini_cheetah_walk", {}),
                                              ("id", "tl_mini_cheetah_walk", {"torque_limits": 1})])
def test_certificate_accepts_oracle_optimum_and_rejects_suboptimal(kind, case, params):
    from quadruped_drake_b200 import load_robot
    g = np.load(GOLD / f"{case}.npz")
    n = 10
    q, v, traj, contact = g["q"][:n], g["v"][:n], g["traj"][:n], g["contact"][:n]
    plant, model = Plant("mini_cheetah"), load_robot("mini_cheetah")
    d = oracle_dynamics(plant, q, v)
    out = oracle_outputs(kind, plant, q, v, traj, contact, **params)
    cert = kkt.certificate(kind, d, model, q, v, traj, contact, out, params)
    assert cert["stationarity"].max() < 1e-8 and cert["dual"].min() > -1e-9 and cert["comp"].max() < 1e-7
    assert cert["eq"].max() < 1e-8 and cert["ineq"].max() < 1e-8
    # a feasible point that is optimal for a DIFFERENT cost (half the body weight): constraints hold, stationarity does not
    other = dict(params)
    other.update({"id_w_body": 5.0} if kind == "id" else {"clf_w_delta": 1.0, "clf_q_foot_p": 400.0})
    sub = oracle_outputs(kind, plant, q, v, traj, contact, **other)
    bad = kkt.certificate(kind, d, model, q, v, traj, contact, sub, params)
    assert bad["eq"].max() < 1e-8 and bad["ineq"].max() < 1e-8
    moved = np.abs(sub.tau - out.tau).max(axis=1) > 1e-4
    assert moved.any() and (bad["stationarity"][moved] > 1e-5).all()


@pytest.mark.parametrize("kind,case,params", [("id", "cfg2_mini_cheetah_stand", {}), ("clf", "mixed_mini_cheetah", {}),
                                              ("id", "tl_mini_cheetah_walk", {"torque_limits": 1})])
def test_device_code_multipliers_certify_the_optimum(built, kind, case, params):
    """The multipliers exported by the device code (compiled for the host, tests/emu) certify its torques."""
    from quadruped_drake_b200 import load_robot
    from quadruped_drake_b200.capi import KINDS, WbcIO, make_params, np_ptr
    emu = C.CDLL(str(Path(__file__).parent / "emu" / "libwbc_emu.so"))
    emu.emu_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]
    g = np.load(GOLD / f"{case}.npz")
    n = 12
    q, v, traj, contact = (np.ascontiguousarray(g[k][:n]) for k in ("q", "v", "traj", "contact"))
    model = load_robot("mini_cheetah")
    ms, pr = model.as_struct(), make_params(**params)
    out = SimpleNamespace(tau=np.zeros((n, 12)), vd=np.zeros((n, 18)), f=np.zeros((n, 4, 3)), qp_info=np.zeros((n, 4)),
                          lam=np.zeros((n, 42)))
    met, st = np.zeros((n, 4)), np.zeros(n, np.int32)
    io = WbcIO(np_ptr(q), np_ptr(v), np_ptr(traj), np_ptr(contact), np_ptr(out.tau), np_ptr(met), np_ptr(st), np_ptr(out.vd),
               np_ptr(out.f), np_ptr(out.qp_info), np_ptr(out.lam))
    assert emu.emu_step(C.byref(ms), C.byref(pr), KINDS[kind], n, C.byref(io)) == 0
    assert (st == 0).all()
    d = oracle_dynamics(Plant("mini_cheetah"), q, v)
    cert = kkt.certificate(kind, d, model, q, v, traj, contact, out, params)
    assert cert["stationarity"].max() < 1e-7 and cert["dual"].min() >= 0.0 and cert["comp"].max() < 1e-6
    assert cert["eq"].max() < 1e-7 and cert["ineq"].max() < 1e-7
    assert np.abs(out.tau - g[f"{kind}_tau"][:n]).max() < 1e-5
