O=gpurun_out; TAG=r02f
timeout 280 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-workloads --no-cpu-baseline --no-other-modes --sustain-s 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dec_tc" -s 2 -c 1 -o $O/${TAG}_prof_train python bench.py --steps 2 --warmup 3 --no-workloads --no-cpu-baseline --no-other-modes --sustain-s 0 > /dev/null 2>&1
timeout 100 python tools/phase_profile.py bridge_p 32768 tc_fp16x3 > $O/${TAG}_phase_bridge_p_tc.log 2>&1
python -c "
import json; d=json.loads(open('$O/${TAG}_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms']); print({k:(v.get('value'),v.get('ms_per_step'),v.get('error')) for k,v in d['workloads'].items()}); print(d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['sustained'])"
