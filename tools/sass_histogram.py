#!/usr/bin/env python
"""SASS instruction histogram of the step kernels of the built library (cuobjdump -sass): the evidence that the bulk-copy /
PDL / FP64 claims of DESIGN.md are in the machine code (UBLKCP = cp.async.bulk, SYNCS = mbarrier, ACQBULK, DFMA, CREDUX =
redux.sync, no HMMA / tensor instructions). Writes profiles/r2_sass_histogram.txt."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "quadruped_drake_b200" / "csrc" / "libwbc_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|<unnamed>::|^void ", "", kern)
        kern = kern.split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
want = [k for k in hist if any(s in k for s in ("wbc_reduce_kernel<0, 1>", "wbc_reduce_kernel<0, 4>", "wbc_solve_kernel<0, false>", "wbc_reduce_pc_kernel", "wbc_plant_kernel", "sample_kernel"))]
lines = ["SASS instruction histogram (static counts, cuobjdump -sass libwbc_b200.so, sm_100a)", ""]
for k in want:
    c = hist[k]
    lines.append(f"{k}: {sum(c.values())} instructions")
    lines.append("  " + ", ".join(f"{op} {n}" for op, n in c.most_common(28)))
    marks = {op: c.get(op, 0) for op in ("UBLKCP", "SYNCS", "ACQBULK", "DFMA", "DMUL", "DADD", "MUFU", "CREDUX", "SHFL", "LDS", "STS", "LDG", "STG", "LDGSTS", "HMMA", "UTCHMMA", "UTMALDG")}
    lines.append("  markers: " + ", ".join(f"{op}={n}" for op, n in marks.items()))
    lines.append("")
txt = "\n".join(lines)
(ROOT / "profiles" / "r2_sass_histogram.txt").write_text(txt)
print(txt)
