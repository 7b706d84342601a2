#!/usr/bin/env python
"""profiles/traffic.json from one ncu --set full capture of a control step (reduce kernel + solve kernel):
DRAM bytes and executed FP64 thread operations per step launch, summed over the kernels of the step.
usage: python tools/ncu_traffic.py gpurun_out/x.ncu-rep instances_per_launch [out.json]"""
import csv, io, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def fnum(s):
    return float(s.replace(",", ""))


def main():
    rep, n = sys.argv[1], int(sys.argv[2])
    outp = Path(sys.argv[3]) if len(sys.argv) > 3 else ROOT / "profiles" / "traffic.json"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(d, name, unit_scale=True):
        i = hdr.index(name)
        v = fnum(d[i])
        u = units[i]
        if unit_scale:
            v *= {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        return v

    seen, kernels = set(), []
    for d in data:
        name = d[hdr.index("Kernel Name")]
        key = name.split("(")[0]
        if key in seen:
            continue          # first launch of each kernel of the step
        seen.add(key)
        cyc = col(d, "smsp__cycles_elapsed.avg", False)
        ops = {op: col(d, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed", False) * cyc for op in ("dfma", "dmul", "dadd")}
        kernels.append({"kernel": key.strip(), "us": col(d, "gpu__time_duration.sum", False),
                        "dram_read": col(d, "dram__bytes_read.sum"), "dram_write": col(d, "dram__bytes_write.sum"),
                        "warp_instructions": col(d, "smsp__inst_executed.sum", False), "fp64_thread_ops": ops})
    tot_ops = {op: sum(k["fp64_thread_ops"][op] for k in kernels) for op in ("dfma", "dmul", "dadd")}
    flops = (2 * tot_ops["dfma"] + tot_ops["dmul"] + tot_ops["dadd"]) / n
    out = {
        "dram_bytes_per_launch": sum(k["dram_read"] + k["dram_write"] for k in kernels),
        "launch": f"one control step = {' + '.join(k['kernel'] for k in kernels)}, {n} instances",
        "source": f"ncu --set full, {Path(rep).name} (dram__bytes_read.sum + dram__bytes_write.sum of the step's kernels)",
        "kernels": kernels,
        "fp64_thread_ops_per_launch": tot_ops,
        "fp64_flops_per_instance_executed": flops,
        "warp_instructions_per_instance": sum(k["warp_instructions"] for k in kernels) / n,
        "fp64_note": "predicated-on thread-level DFMA (x2) + DMUL + DADD of the same capture, per instance: the step's executed FP64 work, "
                     "as opposed to SURVEY 8d's canonical 1.75 MFLOP of the reference-size interior point",
    }
    outp.write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
