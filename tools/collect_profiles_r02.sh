#!/bin/bash
# Round-2 profile set (run on a GPU box through gpurun): test log, bench lines, ncu launch list and full captures,
# phase breakdowns.  Everything lands in gpurun_out/ with the given tag; tools/publish_profiles_r02.py copies the
# summaries into profiles/.
TAG=${1:-r02}
O=gpurun_out
S="--steps 2 --warmup 3 --no-workloads --no-cpu-baseline --no-other-modes --sustain-s 0"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/${TAG}_pytest_gpu.log
timeout 280 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 280 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> /dev/null
timeout 200 python bench.py --scaling strong --steps 5 --warmup 3 --no-workloads --no-cpu-baseline --no-other-modes --sustain-s 0 > $O/${TAG}_bench_1gpu_strong.json 2> /dev/null
timeout 200 python bench.py --workload bridge_s --steps 20 --warmup 5 --no-workloads --no-cpu-baseline --no-other-modes --sustain-s 0 > $O/${TAG}_bench_bridge_s.json 2> /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py $S > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dec_tc|lat_pair|lat_noise|enc_tc|prior_" -s 16 -c 8 -o $O/${TAG}_prof_train python bench.py $S > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"enc_fused|noise_fill|enc_tc_fwd|lat_encode" -s 4 -c 4 -o $O/${TAG}_prof_encode python bench.py --workload bridge_encode --steps 2 --warmup 3 > /dev/null 2>&1
timeout 100 python tools/phase_profile.py bridge_p 32768 tc_fp16x3 > $O/${TAG}_phase_bridge_p_tc.log 2>&1
timeout 100 python tools/phase_profile.py beam_s 32768 tc_fp16x3 > $O/${TAG}_phase_beam_s_tc.log 2>&1
timeout 100 python tools/small_batch_kernel_times.py > $O/${TAG}_small_batch_kernel_times.log 2>&1
timeout 100 python tools/dec_probe.py > $O/${TAG}_dec_probe.log 2>&1
tail -4 $O/${TAG}_pytest_gpu.log
python -c "
import json; d=json.loads(open('$O/${TAG}_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms']); print({k:(v.get('value'),v.get('ms_per_step'),v.get('error')) for k,v in d['workloads'].items()}); print(d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
r=json.loads(open('$O/${TAG}_bench_reference_arm.json').read().strip().splitlines()[-1]); print('ref arm', r['value'], r['cpu_baseline'])"
cat $O/${TAG}_small_batch_kernel_times.log | tail -2
