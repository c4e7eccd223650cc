#!/bin/bash
# Collect the judged artefacts of one round on a GPU box (run under gpurun from the repo root):
#   tools/collect_profiles.sh <tag>      -> gpurun_out/<tag>_*  (copy what should be tracked into profiles/)
# Order matters: every ncu run comes AFTER the same command has exited 0 without ncu.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v Namespace | tail -5 > $out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
timeout 300 python bench.py --workload beam_s --steps 20 --warmup 5 --no-other-modes > $out/${tag}_bench_beam_s.json 2> /dev/null
timeout 300 python bench.py --workload bridge_encode --steps 20 --warmup 5 > $out/${tag}_bench_bridge_encode.json 2> /dev/null
timeout 300 python bench.py --workload ensemble --steps 10 --warmup 3 > $out/${tag}_bench_ensemble.json 2> /dev/null
DPIVAE_BENCH_MEMBERS=16 timeout 300 python bench.py --workload ensemble --steps 10 --warmup 3 > $out/${tag}_bench_ensemble16.json 2> /dev/null
timeout 200 python tools/phase_profile.py bridge_p 32768 tc_fp16x3 > $out/${tag}_phase_bridge_p_tc.log 2>&1
timeout 200 python tools/phase_profile.py beam_s 32768 tc_fp16x3 > $out/${tag}_phase_beam_s_tc.log 2>&1
timeout 200 python tools/small_batch_kernel_times.py > $out/${tag}_small_batch_kernel_times.log 2>&1
# launch list of the same bench command (cold-cache, serialised: shares must agree with the CUDA-event split, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-modes > $out/${tag}_ncu_launches.log 2>&1
python tools/launch_summary.py $out/${tag}_launches.csv > $out/${tag}_launch_shares.txt
# full capture of the dominant kernel (one launch, after warm-up)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dec_tc_kernel -s 6 -c 1 -o $out/${tag}_prof_dec_tc -f \
  python bench.py --steps 3 --warmup 5 --no-cpu-baseline --no-other-modes > $out/${tag}_ncu_dec_tc.log 2>&1
tail -3 $out/${tag}_pytest_gpu.log
python -c "
import json
for n in ('bench','bench_beam_s','bench_bridge_encode','bench_ensemble','bench_ensemble16','bench_reference_arm'):
    try:
        d=json.load(open('$out/${tag}_'+n+'.json')); print(n, d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
    except Exception as e: print(n, 'ERR', e)
"
