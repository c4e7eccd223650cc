#!/bin/bash
# Quick GPU check used between kernel changes (run through gpurun): GPU tests, short bridge_p bench line, small-batch
# per-kernel times, 8-member ensemble.   bash tools/quick_check.sh <tag>
TAG=${1:-q}
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/${TAG}_pytest_gpu.log; tail -3 $O/${TAG}_pytest_gpu.log
for w in bridge_p; do
  timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-workloads --no-cpu-baseline --no-other-modes --sustain-s 0 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
  python -c "
import json; d=json.loads(open('$O/${TAG}_bench_$w.json').read().strip().splitlines()[-1]); print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['elbo'])"
done
timeout 100 python tools/small_batch_kernel_times.py
timeout 100 python tools/phase_profile.py bridge_p 32768 tc_fp16x3 | head -3
timeout 200 python bench.py --workload ensemble --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ensemble', d['value'], d['ms_per_step'])"
