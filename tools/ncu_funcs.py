#!/usr/bin/env python
"""Aggregate an ncu capture by device function of csrc/wbc_device.cuh (see ncu_lines.py).
usage: python tools/ncu_funcs.py gpurun_out/prof.ncu-rep [kernel-substring] [instances]"""
import collections, csv, io, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "wbc_step_kernelILi0E"
ninst = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
src = (ROOT / "quadruped_drake_b200/csrc/wbc_device.cuh").read_text().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template <[^>]*>\s*)?WBC_DEV\s+[\w:<>&\s\*]+?\s(\w+)\(", l) or re.match(r"^inline void (\w+)\(", l)
    if m:
        marks.append((i, m.group(1)))
def func(line):
    name = "?"
    for i, n in marks:
        if i <= line:
            name = n
        else:
            break
    return name
cub = "/tmp/ncu_lines.cubin"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin",
                f"-I{ROOT}/include", str(ROOT / "quadruped_drake_b200/csrc/wbc_api.cu"), "-o", cub], check=True)
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
        off2line[int(m.group(1), 16)] = (cur, m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Address" or r[0].startswith("Kernel"):
        if data:
            break
        continue
    data.append(r)
base = int(data[0][ia], 16)
byf, sf, byop, tot, tots = collections.Counter(), collections.Counter(), collections.Counter(), 0, 0
# attribute intrinsics-header lines to the enclosing function: remember the last wbc_device.cuh function seen
last = "?"
for r in data:
    ent = off2line.get(int(r[ia], 16) - base)
    ln, op = ent if ent else (None, "?")
    if ln and ln[0] == "wbc_device.cuh":
        last = func(ln[1])
    f = last
    byf[f] += int(r[ie]); sf[f] += int(r[isamp]); tot += int(r[ie]); tots += int(r[isamp])
    byop[op.split(".")[0]] += int(r[ie])
print(f"executed {tot} warp-instructions = {tot / ninst:.0f} per instance; {tots} samples")
for f, c in byf.most_common(25):
    print(f"{f:24s} inst {100 * c / tot:6.2f}%  ({c / ninst:7.0f}/inst)  samples {100 * sf[f] / max(tots, 1):6.2f}%")
print("opcode mix:", ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in byop.most_common(18)))
