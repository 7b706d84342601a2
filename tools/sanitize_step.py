#!/usr/bin/env python
"""Small run of every controller kind for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
  compute-sanitizer --tool racecheck python tools/sanitize_step.py
Checks the results against the golden fixtures as well, so a sanitizer-clean run is also a correct one."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from quadruped_drake_b200.controller import BatchedController  # noqa: E402

g = np.load(ROOT / "tests" / "golden" / "mixed_mini_cheetah.npz")
ctl = BatchedController("mini_cheetah", device=0)
for kind, tol in (("id", 1e-5), ("clf", 1e-5), ("pc", 1e-5), ("mptc", 1e-5)):
    out = ctl.step(kind, g["q"], g["v"], g["traj"], g["contact"], debug=True)
    ok = g[f"{kind}_ok"]
    err = np.abs(out.tau - g[f"{kind}_tau"])[ok].max()
    print(kind, "max |tau - golden| = %.3g" % err, "status", np.unique(out.status))
    assert err < tol
ctl2 = BatchedController("mini_cheetah", device=0, torque_limits=1)
out = ctl2.step("id", g["q"], g["v"], g["traj"], g["contact"])
print("torque limits: status", np.unique(out.status))
print("sanitize run ok")

# rows next to the step: trajectory sampler + closed-loop rollout (graph replay) + wire codecs
from quadruped_drake_b200 import planner as pl  # noqa: E402
from quadruped_drake_b200.rollout import rollout  # noqa: E402

Q0 = np.array([1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.3] + [0.0, -0.8, 1.6] * 4)
sampler = pl.TrajectorySampler(ctl, pl.make_gait_plan("mini_cheetah", 0))
n = 70
t0 = np.linspace(0.0, 2.5, n)
r = rollout(ctl, sampler, "id", np.tile(Q0, (n, 1)), np.zeros((n, 18)), t0, 6, 5e-3, log_metrics=True)
print("rollout: status_or", np.unique(r.status_or), "finite", bool(np.isfinite(r.q).all()))
assert np.isfinite(r.q).all()
ref = sampler.sample(t0)
print("sampler rows", np.asarray(ref["traj"] if isinstance(ref, dict) and "traj" in ref else list(ref.values())[0]).shape)
print("sanitize aux ok")

# ground-contact plant (one step and a short closed loop), the device-plan host step and the multi-device call
from quadruped_drake_b200.rollout import plant_step  # noqa: E402
from quadruped_drake_b200.sharding import MultiGpuController  # noqa: E402

rng = np.random.default_rng(0)
q = np.tile(Q0, (n, 1)); q[:, 6] = 0.28; q[:, 7:] += rng.uniform(-0.1, 0.1, (n, 12))
qn, vn, f, st = plant_step(ctl, q, np.zeros((n, 18)), rng.uniform(-3, 3, (n, 12)), 5e-3)
assert np.isfinite(qn).all() and (st == 0).all()
r = rollout(ctl, sampler, "id", np.tile(Q0, (n, 1)), np.zeros((n, 18)), t0, 6, 5e-3, plant=True)
assert np.isfinite(r.q).all()
a = ctl.step_plan("id", sampler, np.tile(Q0, (n, 1)), np.zeros((n, 18)), t0)
assert np.isfinite(a.tau).all()
multi = MultiGpuController("mini_cheetah", devices=[0, 0])
o = multi.step("id", g["q"], g["v"], g["traj"], g["contact"])
assert np.abs(o.tau - g["id_tau"]).max() < 1e-5
multi.close()
print("sanitize plant / plan / multi ok")

# chunked issue of a device step (2-4 reduce -> solve chains on as many streams), the zero-copy host path in two halves and the
# copy-engine-staged host path (24576+ instances): tiled golden instances, bit-identical to the small batch
import torch  # noqa: E402
from quadruped_drake_b200 import capi  # noqa: E402

base = ctl.step("id", g["q"], g["v"], g["traj"], g["contact"])
for nn in (4100, 8200, 24580):
    idx = np.arange(nn) % len(g["q"])
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(g[k][idx])).to(dev) for k in ("q", "v", "traj", "contact")]
    o = ctl.step("id", *t)
    torch.cuda.synchronize()
    assert np.array_equal(o.tau.cpu().numpy(), base.tau[idx]), nn
    hb = [capi.pinned_empty((nn, w), dt) for w, dt in ((19, np.float64), (18, np.float64), (54, np.float64), (4, np.uint8))]
    for dst, k in zip(hb, ("q", "v", "traj", "contact")):
        dst[:] = g[k][idx]
    tau, met, st = capi.pinned_empty((nn, 12)), capi.pinned_empty((nn, 4)), capi.pinned_empty((nn,), np.int32)
    import ctypes as C  # noqa: E402
    io = capi.WbcIO(*(capi.np_ptr(x) for x in hb), capi.np_ptr(tau), capi.np_ptr(met), capi.np_ptr(st), None, None, None)
    assert ctl.lib.wbc_step_host(ctl._h, capi.WBC_CTRL_ID, nn, C.byref(io)) == 0
    assert np.array_equal(tau, base.tau[idx]), nn
print("sanitize chunked / host paths ok")
