"""Copy the artefacts of tools/collect_profiles.sh <tag> from gpurun_out/ into the tracked profiles/ directory
(round-prefixed names), write the ncu summary / hot lines of the dominant kernel and refresh profiles/dec_traffic.json.

    python tools/publish_profiles.py r01g r01
"""
import csv
import json
import os
import shutil
import subprocess
import sys

tag, rnd = sys.argv[1], sys.argv[2]
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
pairs = {"bench.json": "bench_final.json", "bench_reference_arm.json": "bench_reference_arm.json", "bench_beam_s.json": "bench_beam_s.json",
         "bench_bridge_encode.json": "bench_bridge_encode.json", "bench_ensemble.json": "bench_ensemble.json",
         "bench_ensemble16.json": "bench_ensemble16.json", "launches.csv": "launches_final.csv", "launch_shares.txt": "launch_shares_final.txt",
         "phase_bridge_p_tc.log": "phase_bridge_p_tc_final.log", "phase_beam_s_tc.log": "phase_beam_s_tc_final.log",
         "small_batch_kernel_times.log": "small_batch_kernel_times.log", "pytest_gpu.log": "pytest_gpu.log"}
for a, b in pairs.items():
    shutil.copy(os.path.join(G, f"{tag}_{a}"), os.path.join(PR, f"{rnd}_{b}"))
rep = os.path.join(G, f"{tag}_prof_dec_tc.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, d = rows[0], rows[1], dict(zip(rows[0], rows[2]))
f = lambda k: float(d[k].replace(",", "") or 0)
keys = [k for k in hdr if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k]
tot = sum(f(k) for k in keys)
out = [f"ncu --set full --clock-control none, {d['Kernel Name'][:60]}, bridge_p 131072 rows x 16 MC ({tag})"]
out += [f"{100 * f(k) / tot:5.1f}%  {k}" for k in sorted(keys, key=lambda k: -f(k))[:10]]
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
          "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]:
    if k in d:
        out.append(f"{k} {d[k]} {units[hdr.index(k)]}")
open(os.path.join(PR, f"{rnd}_ncu_dec_tc_final_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
rd = f("dram__bytes_read.sum") * mul[units[hdr.index("dram__bytes_read.sum")]]
wr = f("dram__bytes_write.sum") * mul[units[hdr.index("dram__bytes_write.sum")]]
tp = os.path.join(PR, "dec_traffic.json")
t = json.load(open(tp))
t["bridge_p:tc_fp16x3"] = {"dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "total": int(rd + wr),
                           "source": f"ncu --set full, profiles/{rnd}_ncu_dec_tc_final_summary.txt (dec_tc_kernel<0,0,64>, 131072 rows x 16 MC)"}
json.dump(t, open(tp, "w"), indent=1)
