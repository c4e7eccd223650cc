"""Copy the artefacts of tools/collect_profiles_r02.sh <tag> (and tools/run_8gpu_records.sh <tag>) from gpurun_out/ into the
tracked profiles/ directory, write the ncu summaries and refresh profiles/dec_traffic.json (stamped with the commit).

    python tools/publish_profiles_r02.py r02e
"""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
for f in sorted(glob.glob(os.path.join(G, f"{tag}_*"))):
    name = os.path.basename(f)
    if name.endswith((".json", ".csv", ".log")) and os.path.getsize(f) > 0:
        shutil.copy(f, os.path.join(PR, name))
lc = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lc):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), lc], capture_output=True, text=True).stdout
    open(os.path.join(PR, f"{tag}_launch_shares.txt"), "w").write(out)
for kind in ("train", "encode"):
    rep = os.path.join(G, f"{tag}_prof_{kind}.ncu-rep")
    if os.path.exists(rep):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
        open(os.path.join(PR, f"{tag}_ncu_{kind}_summary.txt"), "w").write(out)
# DRAM traffic of the dominant kernel (per launch) for bench.py's roofline.traffic
rep = os.path.join(G, f"{tag}_prof_train.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "dec_tc_kernel<0, 0, 64>" in d.get("Kernel Name", ""):
            f = lambda k: float(d[k].replace(",", "") or 0) * mul[units[hdr.index(k)]]
            rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
            tp = os.path.join(PR, "dec_traffic.json")
            t = json.load(open(tp))
            t["bridge_p:tc_fp16x3"] = {"dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "total": int(rd + wr),
                                       "source": f"ncu --set full, profiles/{tag}_ncu_train_summary.txt (dec_tc_kernel<0,0,64>, 131072 rows x 16 MC)",
                                       "commit": commit}
            json.dump(t, open(tp, "w"), indent=1)
            print("dec_tc traffic", int(rd + wr), "bytes per launch")
            break
print("published", tag, "at", commit)
