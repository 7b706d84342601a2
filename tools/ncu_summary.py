#!/usr/bin/env python
"""Compact per-kernel summary of an ncu report (selected counters + stall reasons).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [instances_per_launch]"""
import csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'smsp__cycles_elapsed.avg',
        'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print("-----", d[hdr.index("Kernel Name")][:70])
        for w in WANT:
            if w in hdr:
                v = d[hdr.index(w)]
                extra = ""
                if n and w in ("smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") or n and "thread_inst_executed_op" in w:
                    try:
                        extra = "   (%.0f per instance)" % (float(v.replace(",", "")) / n)
                    except ValueError:
                        pass
                print("   %-75s %s %s%s" % (w, v, units[hdr.index(w)], extra))
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
                try:
                    if float(d[i]) > 0.15:
                        print("   stall %-30s %s" % (h.split("stalled_")[1].split("_per")[0], d[i]))
                except ValueError:
                    pass


if __name__ == "__main__":
    main()
