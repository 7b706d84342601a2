#!/usr/bin/env python
"""Device-timed throughput of the rows next to the control step: LCM wire codecs, trajectory sampler,
closed-loop rollout. One JSON line per workload, same keys as bench.py (`python bench.py --workload wire|traj|rollout` calls
this). Inputs are resident in HBM and larger than the 126 MB L2 (no flush needed); CUDA events on the launch stream."""
from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _time(fn, steps, warmup):
    import torch
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in evs])


def _line(metric, unit, n, per_ms, bytes_per_unit, workload, steps, warmup, launches, extra=None):
    from bench import measured_peaks
    peak, src = measured_peaks()
    ms = float(per_ms.mean())
    ach = bytes_per_unit * n / (ms * 1e-3) / 1e9
    d = {"metric": metric, "value": n / (ms * 1e-3), "unit": unit, "n_gpus": 1, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms,
         "p50_ms_per_step": float(np.median(per_ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64",
         "data": "synthetic", "config": {"workload": workload, "units_per_launch": n, "l2": "inputs + outputs larger than L2"},
         "gpu_launches": launches,
         "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                      "algorithmic_bytes_per_unit": bytes_per_unit, "peak_source": src}}
    if extra:
        d.update(extra)
    return d


def run(workload, steps=50, warmup=5, n=1 << 20):
    for d in collect(workload, steps, warmup, n):
        print(json.dumps(d))


def step_configs(steps=10, warmup=3):
    """Device-timed control-step throughput of the other BASELINE configs (short runs, inputs resident, L2-sized batches):
    configs[2] anymal_b trot 16384 with the torque box, configs[3] CLF / PC / MPTC at 65536 walk states."""
    import torch
    from quadruped_drake_b200.controller import BatchedController
    from quadruped_drake_b200.synth import generate
    out = {}
    for name, robot, kind, pattern, n, params in (("configs[2] anymal_b trot 16384 ID + torque limits", "anymal_b", "id", "trot", 16384, {"torque_limits": 1}),
                                                  ("configs[3] mini_cheetah CLF walk 65536", "mini_cheetah", "clf", "walk", 65536, {}),
                                                  ("configs[3] mini_cheetah PC walk 65536", "mini_cheetah", "pc", "walk", 65536, {}),
                                                  ("configs[3] mini_cheetah MPTC walk 65536", "mini_cheetah", "mptc", "walk", 65536, {})):
        ctl = BatchedController(robot, device=0, **params)
        q, v, traj, contact = generate(ctl.model, n, 20260120, pattern, ctl.fk)
        dev = torch.device("cuda:0")
        t = [torch.from_numpy(x).to(dev) for x in (q, v, traj)] + [torch.from_numpy(contact).to(dev)]
        o = [torch.empty((n, 12), dtype=torch.float64, device=dev), torch.empty((n, 4), dtype=torch.float64, device=dev),
             torch.empty((n,), dtype=torch.int32, device=dev)]
        io = ctl.make_io(*t, *o)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ctl.time_step(kind, io, n, warmup, stream)
        ms = ctl.time_step(kind, io, n, steps, stream)
        st = o[2].cpu().numpy()
        out[name] = {"value": n / (ms * 1e-3), "unit": "steps/s", "ms_per_step": ms, "instances": n,
                     "solved_fraction": float(((st == 0) | (st == 64)).mean())}
        ctl.close()
    return out


def summary(steps=10, warmup=3):
    """Compact record of the rows next to the step for bench.py's main JSON line (`aux` key): value, unit, HBM fraction."""
    out = {"other_configs": step_configs(steps, warmup)}
    for w in ("wire", "traj", "rollout"):
        for d in collect(w, steps, warmup, 1 << 20, rollout_sizes=((4096, 100),)):
            key = d["config"]["workload"].split(",")[0].split(":")[0]
            if w == "rollout":
                key += " plant=" + str(d.get("plant", 0))
            out[key] = {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                        "hbm_frac": d["roofline"]["frac"] if w != "rollout" else None, "workload": d["config"]["workload"]}
    return out


def collect(workload, steps=50, warmup=5, n=1 << 20, rollout_sizes=((4096, 200), (65536, 50))):
    import torch
    from quadruped_drake_b200 import planner as pl
    from quadruped_drake_b200.controller import BatchedController
    from quadruped_drake_b200.wire import WireCodec
    ctl = BatchedController("mini_cheetah", device=0)
    out = []
    rng = np.random.default_rng(0)
    if workload == "wire":
        w = WireCodec(ctl)
        traj, f = rng.normal(0, 3, (n, 54)), rng.normal(0, 40, (n, 12))
        contact = rng.integers(0, 2, (n, 4)).astype(np.uint8)
        tt = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
        msgs = w.encode_trunk_state(tt(np.arange(n) * 1e-3), tt(np.zeros(n, np.uint8)), tt(traj), tt(contact), tt(f))   # synthetic wire bytes
        d = w.decode_trunk_state(msgs)
        assert int(d["status"].max().item()) == 0 and torch.equal(d["traj"], tt(traj))
        l0 = ctl.launches
        per = _time(lambda: w.decode_trunk_state(msgs), steps, warmup)
        out.append(_line("trunk_state_t messages decoded/sec", "messages/s", n, per, 549 + 545, "wbc_lcm_decode_trunk_state, 1Mi messages", steps, warmup, ctl.launches - l0))
        per = _time(lambda: w.encode_trunk_state(d["timestamp"], d["finished"], d["traj"], d["contact"], d["f"]), steps, warmup)
        out.append(_line("trunk_state_t messages encoded/sec", "messages/s", n, per, 541 + 549, "wbc_lcm_encode_trunk_state, 1Mi messages", steps, warmup, steps))
        tau = torch.from_numpy(rng.normal(0, 10, (n, 12))).cuda()
        rm, _ = w.encode_robot_state(None, None, tau, tau_in_actuator_order=True)
        per = _time(lambda: w.encode_robot_state(None, None, tau, tau_in_actuator_order=True), steps, warmup)
        out.append(_line("robot_state_control_lcmt torque messages encoded/sec", "messages/s", n, per, 96 + 204 + 4, "wbc_lcm_encode_robot_state (torques only), 1Mi messages", steps, warmup, steps))
        per = _time(lambda: w.decode_robot_state(rm), steps, warmup)
        out.append(_line("robot_state_control_lcmt messages decoded/sec", "messages/s", n, per, 204 + 49 * 8 + 4, "wbc_lcm_decode_robot_state, 1Mi messages", steps, warmup, steps))
    elif workload == "traj":
        plans = [pl.make_gait_plan("mini_cheetah", c) for c in ("walk", "trot", "pace", "bound")]
        s = pl.TrajectorySampler(ctl, plans)
        t = torch.from_numpy(rng.uniform(0, 5, n)).cuda()
        pi = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32)).cuda()
        per = _time(lambda: s.sample(t, pi), steps, warmup)
        out.append(_line("trunk-trajectory samples/sec", "samples/s", n, per, 12 + 432 + 4 + 8 + 4, "wbc_sample_trajectory, 4 gait plans, 1Mi (plan, t) pairs", steps, warmup, steps))
    elif workload == "rollout":
        from quadruped_drake_b200.rollout import Q0_MINI_CHEETAH as Q0, rollout
        for nr, k, plant in [(a, b, p) for a, b in rollout_sizes for p in (False, True)]:
            bh = Q0[6] - ctl.dynamics(Q0[None], np.zeros((1, 18)))["p_feet"][0, :, 2].mean()
            plans = [pl.make_motion_plan("mini_cheetah", m, 6.0, base_height=bh, phase=ph) for m in ("orientation", "heave", "raise_foot")
                     for ph in np.linspace(0, 2 * np.pi, 8, endpoint=False)]
            s = pl.TrajectorySampler(ctl, plans)
            q0 = np.tile(Q0, (nr, 1)); q0[:, 6] = bh
            pi = torch.from_numpy(rng.integers(0, len(plans), nr).astype(np.int32)).cuda()
            res = {}
            for graph in (True, False):
                def once():
                    q, v, t = (torch.from_numpy(x).cuda() for x in (q0, np.zeros((nr, 18)), np.zeros(nr)))
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    r = rollout(ctl, s, "id", q, v, t, k, 5e-3, plan_index=pi, use_graph=graph, plant=plant)
                    e1.record()
                    torch.cuda.synchronize()
                    assert int(r.status_or.max().item()) == 0
                    return e0.elapsed_time(e1)
                with torch.cuda.stream(torch.cuda.Stream()):
                    for _ in range(2):
                        once()
                    res[graph] = np.array([once() for _ in range(5)])
            per = res[True]
            what = "sample + ID-QP + ground-contact plant step" if plant else "sample + ID-QP + integrate"
            d = _line(f"closed-loop robot control steps/sec ({what})", "robot-steps/s", nr * k, per, 860 + 2 * 444 + 2 * 296 + 144,
                      f"wbc_rollout: {nr} robots x {k} steps, reference test motions, CUDA-graph replay, "
                      + ("torques applied to the simulated robot on the ground" if plant else "QP accelerations integrated"), 5, 2, 4 * k * 5,
                      {"without_graph_robot_steps_per_s": nr * k / (float(res[False].mean()) * 1e-3), "plant": int(plant)})
            out.append(d)
    else:
        raise SystemExit(f"unknown workload {workload}")
    ctl.close()
    return out


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "wire")
