#!/usr/bin/env python
"""Fit the executed-FP64-work model bench.py reports against (profiles/r2_flop_model.json) from two ncu --set full captures
of the step kernels at 4096 instances with different mean active-set iterations.

  python tools/ncu_flops.py A.ncu-rep itersA B.ncu-rep itersB [n_instances]

Per kernel: thread-level predicated-on DFMA x 2 + DMUL + DADD (smsp__sass_thread_inst_executed_op_*), per instance.
reduce_flops = mean of the two captures; solve = base + per_iteration x iterations through the two points. Also records
the DRAM bytes of capture A (the `roofline.traffic` figure: ncu flushes caches between kernels, so this includes the
hand-over record that stays in L2 in a live run) and the utilisation of the most loaded units (the binding resource)."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = {}
    for d in data:
        name = d[hdr.index("Kernel Name")]
        key = "reduce" if "reduce" in name else ("solve" if "solve" in name else None)
        if key is None or key in res:
            continue
        f = lambda m: float(d[hdr.index(m)].replace(",", ""))  # noqa: E731
        byt = lambda m: f(m) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(units[hdr.index(m)], 1.0)  # noqa: E731
        cyc = f("smsp__cycles_elapsed.avg")
        op = lambda o: f(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed") * cyc  # noqa: E731
        res[key] = {"flops": 2 * op("dfma") + op("dmul") + op("dadd"),
                    "dram": byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum"), "us": f("gpu__time_duration.sum"),
                    "inst": f("smsp__inst_executed.sum"),
                    "lsu_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                    "kernel": name[:60]}
    return res


def main():
    repA, itA, repB, itB = sys.argv[1], float(sys.argv[2]), sys.argv[3], float(sys.argv[4])
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
    A, B = load(repA), load(repB)
    per = (A["solve"]["flops"] - B["solve"]["flops"]) / n / (itA - itB)
    base = A["solve"]["flops"] / n - per * itA
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    model = {
        "reduce_flops": 0.5 * (A["reduce"]["flops"] + B["reduce"]["flops"]) / n,
        "solve_flops_base": base, "solve_flops_per_iteration": per,
        "dram_bytes_per_instance_4096": (A["reduce"]["dram"] + A["solve"]["dram"]) / n,
        "warp_instructions_per_instance": {"reduce": A["reduce"]["inst"] / n, "solve": A["solve"]["inst"] / n, "at_iterations": itA},
        "binding_resource": {k: {m: A[k][m] for m in ("lsu_wavefronts_pct", "issue_active_pct", "fp64_pipe_pct", "us")} for k in ("reduce", "solve")},
        "source": f"static, from {Path(repA).name} ({itA} iterations) and {Path(repB).name} ({itB} iterations) @ {commit}, {n} instances, "
                  "ncu --set full --clock-control none; thread-level DFMA x2 + DMUL + DADD",
    }
    (ROOT / "profiles" / "r2_flop_model.json").write_text(json.dumps(model, indent=1) + "\n")
    print(json.dumps(model, indent=1))


if __name__ == "__main__":
    main()
