#!/usr/bin/env python
"""Benchmark of the batched whole-body QP control step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (dynamics + ID-QP: the reduce kernel and the solve kernel, back to back on one
stream) over one batch of synthetic states.
Workload at N = 1: BASELINE.json configs[1] - mini_cheetah ID-QP, 4096 random states per launch, all
four feet in stance. With N > 1 (torchrun, one rank per GPU) every rank runs the same
batch size on its own shard of instances: weak scaling, no collective on the data path.

Printed JSON (rank 0): value = control steps/s with inputs resident in HBM (CUDA events on the launch
stream, L2 flushed between timed launches, max over ranks); e2e = the same through `wbc_step_host` with
pinned host buffers (H2D + kernel + D2H inside the timed region); roofline / cpu_baseline as the
contract asks. `--impl reference` times the CPU restatement of the reference path (oracle/, Python or
C port) on the host cores instead - pydrake/OSQP themselves are not installable.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ROBOT, PATTERN, BATCH, SEED = "mini_cheetah", "stand", 4096, 20260119
METRIC = "whole-body QP control steps/sec"
ALGO_BYTES_PER_STEP = 860          # q,v 296 + traj 432 + contact 4 in; tau 96 + metrics 32 out
CANON_FLOPS_PER_STEP = {0: 0.56e6, 1: 0.78e6, 2: 1.05e6, 3: 1.37e6, 4: 1.75e6}   # SURVEY 8d F(nc) @ 12 IPM iterations
WORKLOAD = "mini_cheetah ID-QP controller step, 4096 random synthetic states (BASELINE configs[1])"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML every ~2 ms from a thread
    (the nvidia-smi recipe of B200_PROFILING.md samples too coarsely for a region of tens of milliseconds)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz, self.err = index, [], set(), None, None
        self._stop = threading.Event()
        self._th = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indices follow PCI order like CUDA_VISIBLE_DEVICES-less CUDA; honour the mask if there is one
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(self.h))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                break
            time.sleep(0.002)

    def stop(self):
        self._stop.set()
        if self._th is not None:
            self._th.join(timeout=2.0)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "NVML polled every 2 ms during the timed region"}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    robot, q, v, traj, contact = args
    from oracle.controllers import IDController, traj_to_dict
    ctl = IDController(robot)
    t0 = time.perf_counter()
    for i in range(len(q)):
        ctl.control_law(q[i], v[i], traj_to_dict(traj[i], contact[i]))
    return time.perf_counter() - t0


def cpu_port_throughput(q, v, traj, contact, budget_s=20.0):
    """Oracle (CPU restatement of the reference path) on all host cores over a bounded sample."""
    cport = ROOT / "oracle" / "_build" / "liboracle_c.so"
    if cport.exists():
        from oracle.cport import time_id_steps
        return time_id_steps(ROBOT, q, v, traj, contact, budget_s)
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    _cpu_worker((ROBOT, q[:2], v[:2], traj[:2], contact[:2]))
    per = (time.perf_counter() - t0) / 2
    per_core = max(2, min(len(q) // cores, int(budget_s / max(per, 1e-3))))
    chunks = [(ROBOT, q[c * per_core:(c + 1) * per_core], v[c * per_core:(c + 1) * per_core],
               traj[c * per_core:(c + 1) * per_core], contact[c * per_core:(c + 1) * per_core]) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, chunks)
    wall = time.perf_counter() - t0
    done = per_core * cores
    return {"value": done / wall, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"{done} of the {len(q)} instances of one batch, numpy oracle (oracle/controllers.py), one process per core"}


def host_inputs(n, seed):
    """Synthetic batch on the host. Forward kinematics for the trajectory targets comes from the oracle here
    only when no GPU is in use (reference arm); the GPU arm uses wbc_dynamics."""
    from quadruped_drake_b200 import load_robot
    from quadruped_drake_b200.synth import generate
    return load_robot(ROBOT), generate


def run_reference(args):
    """Reference arm: the CPU restatement of the reference path (C twin of the oracle when built, else the numpy
    oracle) on all host cores, same config / metric / unit; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    g.build_oracle()
    from quadruped_drake_b200 import load_robot
    from quadruped_drake_b200.synth import generate
    model = load_robot(ROBOT)
    cport = (ROOT / "oracle" / "_build" / "liboracle_c.so").exists()
    if cport:
        from oracle.cport import fk as cfk
        n, fkc = BATCH, cfk(ROBOT)
    else:
        sys.path.insert(0, str(ROOT / "tests"))
        from conftest import oracle_fk
        from oracle.dynamics import Plant
        n, fkc = 64 * (os.cpu_count() or 1), oracle_fk(Plant(ROBOT))
    q, v, traj, contact = generate(model, n, SEED, PATTERN, fkc)
    vals = []
    per_step = max(2.0, min(20.0, 120.0 / (args.warmup + args.steps)))
    for s in range(args.warmup + args.steps):
        r = cpu_port_throughput(q, v, traj, contact, budget_s=per_step)
        if s >= args.warmup:
            vals.append(r)
    val = float(np.mean([r["value"] for r in vals]))
    base = dict(vals[-1])
    base["value"] = val
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * BATCH / val, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "robot": ROBOT, "controller": "ID", "contact_pattern": PATTERN,
                       "instances_per_step": BATCH,
                       "note": "CPU restatement of the reference path on all host cores (pydrake + OSQP are not installable "
                               "here, DESIGN.md 5); each step is a bounded sample of the 4096-instance batch"},
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the controller path has no CPU fallback (use --impl reference for the CPU arm)")
    if rank == 0:
        g.build()
    if world > 1:
        # NCCL announces its version on stdout when the first communicator comes up; stdout carries only the JSON line, so
        # it is pointed at stderr for the rendezvous
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from quadruped_drake_b200 import capi
    from quadruped_drake_b200.controller import BatchedController, measure_fp64_peak
    from quadruped_drake_b200.synth import generate

    robot = args.robot
    kind = capi.KINDS[args.controller]
    ctl = BatchedController(robot, device=local, **({"torque_limits": 1} if args.torque_limits else {}))
    n = args.batch
    q, v, traj, contact = generate(ctl.model, n, SEED + 1000 * rank, args.pattern, ctl.fk)   # each rank its own shard
    tq, tv, tt = (torch.from_numpy(x).to(dev) for x in (q, v, traj))
    tc = torch.from_numpy(contact).to(dev)
    tau = torch.empty((n, 12), dtype=torch.float64, device=dev)
    met = torch.empty((n, 4), dtype=torch.float64, device=dev)
    st = torch.empty((n,), dtype=torch.int32, device=dev)
    qi = torch.empty((n, 4), dtype=torch.float64, device=dev)
    io = ctl.make_io(tq, tv, tt, tc, tau, met, st, None, None, qi)
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)       # 256 MB > 126 MB L2
    import ctypes as C

    def launch():
        rc = ctl.lib.wbc_step(ctl._h, kind, n, C.byref(io), C.c_void_p(stream.cuda_stream))
        if rc:
            raise RuntimeError(ctl.lib.wbc_last_error(ctl._h).decode())

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        launch()
    torch.cuda.synchronize()
    status = st.cpu().numpy()
    iters = float(qi[:, 3].mean().item())
    bad = status != 0
    if args.controller in ("pc", "mptc"):
        bad &= status != capi.ST_UNSUPPORTED          # full-flight instances: the reference PC / MPTC raise there too
    if bad.any():
        raise SystemExit(f"bench.py: {int(bad.sum())} instances returned a non-zero status")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctl.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in evs:
        flush.zero_()                       # L2 flush between timed launches (not inside the event pair)
        e0.record(stream)
        launch()
        e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    gpu_launches = ctl.launches - launches0
    per = np.array([e0.elapsed_time(e1) for e0, e1 in evs])     # ms per launch
    total_ms = float(per.sum())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel share of the step (events between the two kernels; separate short run, same inputs)
    ms_r, ms_s = C.c_double(), C.c_double()
    shares = None
    if n <= 262144 and ctl.lib.wbc_profile_step(ctl._h, kind, n, C.byref(io), 20, C.c_void_p(stream.cuda_stream), C.byref(ms_r), C.byref(ms_s)) == 0:
        shares = {"reduce_kernel_ms": ms_r.value, "solve_kernel_ms": ms_s.value,
                  "solve_share": ms_s.value / max(ms_r.value + ms_s.value, 1e-12)}

    # ---- end to end through the host API: pinned host buffers, H2D + kernel + D2H every step
    hq, hv, ht = capi.pinned_empty((n, 19)), capi.pinned_empty((n, 18)), capi.pinned_empty((n, 54))
    hc = capi.pinned_empty((n, 4), np.uint8)
    hq[:], hv[:], ht[:], hc[:] = q, v, traj, contact
    htau, hmet, hst = capi.pinned_empty((n, 12)), capi.pinned_empty((n, 4)), capi.pinned_empty((n,), np.int32)
    hio = capi.WbcIO(capi.np_ptr(hq), capi.np_ptr(hv), capi.np_ptr(ht), capi.np_ptr(hc), capi.np_ptr(htau), capi.np_ptr(hmet),
                     capi.np_ptr(hst), None, None, None)
    for _ in range(3):
        ctl.lib.wbc_step_host(ctl._h, kind, n, C.byref(hio))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = ctl.lib.wbc_step_host(ctl._h, kind, n, C.byref(hio))
        assert rc == 0
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert np.array_equal(htau, tau.cpu().numpy()), "host and device entry points disagree"
    h2d = n * (19 + 18 + 54) * 8 + n * 4
    d2h = n * (12 + 4) * 8 + n * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * n * args.steps / (total_ms * 1e-3)
    kernel_ms = float(per.mean())
    hbm_peak, peak_src = measured_peaks()
    ach_gbs = ALGO_BYTES_PER_STEP * n / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
    fp64_peak = measure_fp64_peak(local)
    canon = CANON_FLOPS_PER_STEP[4]
    exec_flops = json.loads(tp.read_text()).get("fp64_flops_per_instance_executed") if tp.exists() else None
    ach_tf = (exec_flops or 0.0) * n / (kernel_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "p50_ms_per_step": float(np.median(per)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "robot": robot, "controller": args.controller.upper(), "contact_pattern": args.pattern,
                   "instances_per_step_per_gpu": n, "l2": "flushed between timed launches (256 MB memset outside the event pair)",
                   "mean_active_set_iterations": iters, "tie_break_reg_f": 1e-6},
        "e2e": {"value": world * n * args.steps / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "wbc_step_host on page-locked host buffers, inputs and outputs cross the host link inside the timed region: the kernels "
                        "read / write host memory directly (zero-copy, below 131072 instances per call) or the batch goes through the "
                        "65536-instance two-stream copy / compute pipeline (above; also the path of pageable buffers)"},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernels": shares,
                     "note": "860 B/step of algorithmic I/O over the whole step (reduce + solve kernel; the dominant one is the solve kernel, "
                             "share in `kernels`): this path is bound by FP64 dependent-instruction latency and the shared-memory pipe, not by "
                             "HBM (see roofline_fp64)"},
        "roofline_fp64": {"achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                          "flops_per_step": exec_flops,
                          "basis": "EXECUTED FP64 work: thread-level DFMA x2 + DMUL + DADD per instance from the committed ncu capture of "
                                   "the headline workload (profiles/traffic.json) x measured steps/s; both kernels are bound by dependent-"
                                   "instruction latency and shared-memory wavefronts at 16 / 28 resident warps per SM, not by the FP64 pipe "
                                   "(DESIGN.md 3)",
                          "peak_source": "DFMA loop measured in this run (wbc_measure_fp64_peak)",
                          "canonical": {"flops_per_step": canon, "tflops_equivalent": canon * n / (kernel_ms * 1e-3) / 1e12,
                                        "note": "SURVEY 8d F(nc=4): 12 interior-point iterations on the reference-size KKT system. The "
                                                "null-space + active-set kernel reaches the exact optimum with ~13x fewer flops, so this "
                                                "equivalent rate can exceed the hardware peak"}},
    }
    if n != BATCH or args.pattern != PATTERN or robot != ROBOT or args.controller != "id" or args.torque_limits:
        line["config"]["workload"] = (f"EXPERIMENT (not the BASELINE bench config): {robot} {args.controller.upper()}-QP, {n} instances/launch, "
                                      f"pattern {args.pattern}, torque limits {'on' if args.torque_limits else 'off'}")
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_port_throughput(q[:4096], v[:4096], traj[:4096], contact[:4096])
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="instances per launch per GPU (default: the BASELINE config, 4096)")
    ap.add_argument("--pattern", default=PATTERN, choices=["stand", "trot", "walk", "mixed"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (experiments only)")
    ap.add_argument("--robot", default=ROBOT, choices=["mini_cheetah", "anymal_b"], help="experiments only")
    ap.add_argument("--controller", default="id", choices=["id", "clf", "pc", "mptc"], help="experiments only")
    ap.add_argument("--torque-limits", action="store_true", help="experiments only: |tau| <= effort rows (BASELINE configs[2])")
    ap.add_argument("--workload", default="step", choices=["step", "wire", "traj", "rollout"],
                    help="step = the BASELINE metric (default); wire / traj / rollout = the rows next to it (tools/bench_aux.py)")
    args = ap.parse_args()
    if args.workload != "step":
        import __graft_entry__ as g
        g.build()
        sys.path.insert(0, str(ROOT / "tools"))
        import bench_aux
        bench_aux.run(args.workload, steps=min(args.steps, 50), warmup=args.warmup)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
