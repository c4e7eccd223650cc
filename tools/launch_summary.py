"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, average and total time per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv [name-filter]
"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
flt = sys.argv[2] if len(sys.argv) > 2 else "dpv::"
t = defaultdict(list)
for r in rows[1:]:
    if flt in r[ik]:
        t[r[ik].split("(")[0][:60]].append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in t.values())
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print(f"{100 * sum(v) / tot:5.1f}%  {len(v):4d} launches  avg {sum(v) / len(v) / 1000:9.1f} us  {k}")
