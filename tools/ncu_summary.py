"""Stall breakdown and key counters of every kernel in an ncu report (`ncu --set full`).

    python tools/ncu_summary.py <report.ncu-rep> [> profiles/rNN_ncu_<kernel>_summary.txt]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_pipe_xu.sum", "local_load_sectors" ]
for r in rows[2:]:
    if not r:
        continue
    d = dict(zip(hdr, r))
    f = lambda k: float(d[k].replace(",", "") or 0)
    keys = [k for k in hdr if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k]
    tot = sum(f(k) for k in keys) or 1.0
    print(f"== {d['Kernel Name'][:90]}  (ncu --set full --clock-control none)")
    for k in sorted(keys, key=lambda k: -f(k))[:9]:
        print(f"  {100 * f(k) / tot:5.1f}%  {k}")
    for k in KEYS:
        if k in d:
            print(f"  {k} {d[k]} {units[hdr.index(k)]}")
