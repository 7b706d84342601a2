#!/usr/bin/env python
"""Executed instructions per source line + opcode mix of one kernel of an ncu capture, normalised per instance.
usage: python tools/ncu_lines2.py rep kernel-substring instances [top]   (NCU_LINES_FLAGS: extra nvcc flags of the captured build)"""
import collections, csv, io, os, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
rep, kern, N = sys.argv[1], sys.argv[2], int(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cub = "/tmp/ncu_lines2.cubin"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin", f"-I{ROOT}/include",
                *os.environ.get("NCU_LINES_FLAGS", "").split(), str(ROOT / "quadruped_drake_b200/csrc/wbc_api.cu"), "-o", cub], check=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
cur, off2line, off2op = None, {}, {}
for l in sass[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off = int(m.group(1), 16); off2line[off] = cur
        op = m.group(2).split(); o = op[1] if op[0].startswith('@') else op[0]
        off2op[off] = o.split('.')[0]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the source page lists the captured kernels one after another; take the first block of the kernel asked for
want = {"reduce": "wbc_reduce_kernel", "solve": "wbc_solve_kernel", "reduce_pc": "wbc_reduce_pc_kernel"}.get(os.environ.get("NCU_KERNEL", ""), None)
blocks, curb, hdr, name = [], None, None, ""
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1]; curb = None; continue
    if r and r[0] == "Address":
        hdr = r; curb = []; blocks.append((name, curb)); continue
    if curb is not None and hdr and len(r) >= len(hdr) - 1:
        curb.append(r)
cands = [b for nm, b in blocks if (want is None or want in nm)]
blk = min(cands, key=lambda b: abs(len(b) - len(off2line)))
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(blk[0][ia], 16)
byline, byop, samp, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
lineops = collections.defaultdict(collections.Counter)
for r in blk:
    off = int(r[ia], 16) - base; ln = off2line.get(off); n = int(r[ie])
    byline[ln] += n; byop[off2op.get(off)] += n; tot += n; samp[ln] += int(r[isamp]); lineops[ln][off2op.get(off)] += n
src = (ROOT / "quadruped_drake_b200/csrc/wbc_device.cuh").read_text().split("\n")
print(f"SASS {len(off2line)} (block {len(blk)}), executed per instance {tot / N:.0f}, samples {sum(samp.values())}")
print("opmix", [(k, round(v / N)) for k, v in byop.most_common(24)])
ts = max(sum(samp.values()), 1)
for ln, c in sorted(byline.items(), key=lambda kv: -kv[1])[:top]:
    txt = src[ln[1] - 1].strip()[:90] if ln and ln[0] == "wbc_device.cuh" else str(ln)
    print(f"{c / N:7.0f} {100 * samp[ln] / ts:5.1f}% {ln[1] if ln else 0:5d} {txt}   {[(k, round(v / N)) for k, v in lineops[ln].most_common(3)]}")
