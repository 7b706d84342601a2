#!/usr/bin/env python
"""Attribute executed SASS instructions / stall samples of an ncu capture to source lines.

usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [top]
Recompiles the current csrc/ to a cubin with -lineinfo, maps SASS offsets to lines with nvdisasm and
joins that with `ncu --page source --csv` (per-instruction counters of the first matching kernel).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    rep = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 else "wbc_step_kernelILi0E"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cub = "/tmp/ncu_lines.cubin"
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin",
                    f"-I{ROOT}/include", *os.environ.get("NCU_LINES_FLAGS", "").split(),   # e.g. -DWBC_SOLVE_CTAS=4 for a variant build
                    str(ROOT / "quadruped_drake_b200/csrc/wbc_api.cu"), "-o", cub], check=True)
    sass = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
    cur, off2line = None, {}
    for l in sass[start + 1:]:
        if l.startswith("//-----"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] in ("Address",) or r[0].startswith("Kernel"):
            if data:
                break
            continue
        data.append(r)
    base = int(data[0][ia], 16)
    byline, samp, tot, tots = collections.Counter(), collections.Counter(), 0, 0
    for r in data:
        ln = off2line.get(int(r[ia], 16) - base)
        byline[ln] += int(r[ie]); samp[ln] += int(r[isamp]); tot += int(r[ie]); tots += int(r[isamp])
    src = (ROOT / "quadruped_drake_b200/csrc/wbc_device.cuh").read_text().split("\n")
    print(f"SASS instructions {len(off2line)}, executed {tot}, samples {tots}")
    print(" inst%  samp%  line  source")
    for ln, c in sorted(byline.items(), key=lambda kv: -samp[kv[0]])[:top]:
        txt = src[ln[1] - 1].strip()[:100] if ln and ln[0] == "wbc_device.cuh" else str(ln)
        print(f"{100 * c / tot:6.2f} {100 * samp[ln] / max(tots, 1):6.2f}  {ln[1] if ln else 0:5d} {txt}")


if __name__ == "__main__":
    main()
