#!/bin/bash
# Same-box A/B of the in-tree library against tools/ab/libold.so (box-to-box variation is ~2 %: only same-box pairs count).
#   bash tools/ab_bench.sh [workload ...]
for w in ${@:-bridge_p}; do for i in 1 2; do for e in "A=new" "DPIVAE_B200_LIB=/root/repo/tools/ab/libold.so"; do
  env $e timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-workloads --no-cpu-baseline --no-other-modes --sustain-s 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w [${e%%=*}]', round(d['value']/1e6,2), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['kernel_ms'].items()}, d['elbo'])"
done; done; done
