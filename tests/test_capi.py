"""The C-ABI library loads on a CPU-only box and exports every symbol include/wbc.h declares."""
import ctypes as C
import re
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def header_functions():
    text = (ROOT / "include" / "wbc.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wbc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(built):
    from quadruped_drake_b200 import capi
    lib = capi.load_library()
    names = header_functions()
    assert "wbc_step_id" in names and "wbc_create" in names and len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in wbc.h but not exported"
    assert set(capi.EXPORTED_SYMBOLS) == set(names)


def test_struct_sizes_match_header(built):
    from quadruped_drake_b200.capi import WbcIO, WbcParams
    from quadruped_drake_b200.model import WbcModelStruct
    assert C.sizeof(WbcModelStruct) == 8 * (13 + 39 + 78 + 36 + 36 + 12 + 12 + 3) + 4 * 24
    assert C.sizeof(WbcParams) == 8 * 29 + 8 + 8 * 3 + 8 * 12
    assert C.sizeof(WbcIO) == 88          # 11 pointers (q v traj contact tau metrics status vd f qp_info lam)
    from quadruped_drake_b200.capi import WbcPlantOpts, WbcRolloutIO, WbcRolloutOpts
    assert C.sizeof(WbcPlantOpts) == 24 and C.sizeof(WbcRolloutOpts) == 40 and C.sizeof(WbcRolloutIO) == 72
    lib = __import__("quadruped_drake_b200.capi", fromlist=["load_library"]).load_library()
    o = WbcPlantOpts()
    assert lib.wbc_default_plant_opts(C.byref(o)) == 0 and (o.mu, o.erp, o.iters) == (1.0, 0.2, 30)     # simulate.py:44-46: friction 1.0


def test_default_params_match_reference_constants(built):
    from quadruped_drake_b200 import capi
    lib = capi.load_library()
    p = capi.WbcParams()
    assert lib.wbc_default_params(C.byref(p)) == 0
    q = capi.make_params()
    for name, _ in capi.WbcParams._fields_:
        a, b = getattr(p, name), getattr(q, name)
        assert (list(a) == list(b)) if name == "pd_q_nom" else (a == b), name
    assert (p.mu, p.contact_damping, p.id_kp_body_p, p.id_w_body, p.clf_w_delta, p.pc_kp_foot) == (0.7, 100, 500, 10, 1000, 200)


def test_create_fails_loudly_without_gpu_or_bad_args(built):
    """No silent CPU fallback: without a device wbc_create returns an error with a message."""
    import torch
    from quadruped_drake_b200 import capi, load_robot
    lib = capi.load_library()
    ms = load_robot("mini_cheetah").as_struct()
    h = C.c_void_p()
    bad = capi.make_params(reg_f=0.0)
    assert lib.wbc_create(C.byref(ms), C.byref(bad), 0, C.byref(h)) == 1
    assert b"reg_f" in lib.wbc_last_error(h)
    lib.wbc_destroy(h)
    if not torch.cuda.is_available():
        h = C.c_void_p()
        pr = capi.make_params()
        rc = lib.wbc_create(C.byref(ms), C.byref(pr), 0, C.byref(h))
        assert rc == 2 and len(lib.wbc_last_error(h)) > 0
        lib.wbc_destroy(h)


def test_model_tables():
    from quadruped_drake_b200 import load_robot
    for robot, mass in (("mini_cheetah", 8.252), ("anymal_b", 30.421396462)):
        m = load_robot(robot)
        assert abs(m.total_mass - mass) < 1e-9
        B = m.actuation_matrix()
        assert B.shape == (18, 12) and (B.sum(0) == 1).all() and (B[:6] == 0).all()
        bf = load_robot(robot, "breadth_first")
        assert sorted(bf.v_index.tolist()) == list(range(6, 18)) and bf.v_index[1] == 10
    a = load_robot("anymal_b")
    assert np.allclose(a.foot_xyz[0], [0.1, -0.02, -0.32125])        # adapter + foot offsets merged
    assert abs(a.mass[3] - (0.207204302 + 0.140170767)) < 1e-12      # shank + welded adapter
