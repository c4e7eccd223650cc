"""Attribute ncu per-SASS-instruction samples to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <kernel cubin> [top]

Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line markers of the same cubin by
instruction index, then prints samples per (file, line), with the inlined-at call line when the
instruction comes from common.cuh."""
import csv
import re
import subprocess
import sys
from collections import Counter, defaultdict

rep, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
import os
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []), capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r_ in rows[2:]:
    if r_ and r_[0] == "Kernel Name":   # several captured launches: keep the first
        break
    data.append(r_)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []  # per instruction: (file, line, inlined_at)
cur = ("?", 0, "")
inl = ""
in_kernel = False
kname = sys.argv[4] if len(sys.argv) > 4 else rows[0][1].split("(")[0].split("::")[-1].rstrip("<")
for l in dis:
    if l.startswith("\t.section") or l.startswith(".section"):
        in_kernel = kname in l and ".text." in l
    if not in_kernel:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), f"{m.group(3).split('/')[-1]}:{m.group(4)}" if m.group(3) else "")
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print("sass instr (ncu, nvdisasm):", len(data), len(lines))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = Counter()
agg_st = defaultdict(Counter)
inst = Counter()
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else ("?", 0, "")
    s = int(r[ix["# Samples"]] or 0)
    agg[key] += s
    inst[key] += int(r[ix["Instructions Executed"]] or 0)
    for st in stalls:
        v = int(r[ix[st]] or 0)
        if v:
            agg_st[key][st[6:]] += v
tot = sum(agg.values())
print("total samples", tot)
for key, v in agg.most_common(top):
    st = ", ".join(f"{k}:{c}" for k, c in agg_st[key].most_common(3))
    print(f"{100*v/tot:5.1f}%  {key[0]}:{key[1]:<4d} {('<- ' + key[2]) if key[2] else '':22s} inst={inst[key]:>11d}  {st}")
