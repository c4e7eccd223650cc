timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -2
for i in 1 2; do for e in "A=fused" "DPIVAE_NO_FUSED_ADAM=1"; do
  env $e timeout 200 python bench.py --steps 30 --warmup 5 --no-workloads --no-cpu-baseline --no-other-modes --sustain-s 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$e]', round(d['value']/1e6,2), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['kernel_ms'].items()}, d['elbo'], d['gpu_launches'])"
done; done
for e in "A=fused" "DPIVAE_NO_FUSED_ADAM=1"; do env $e timeout 100 python tools/small_batch_kernel_times.py | head -1; env $e timeout 200 python bench.py --workload ensemble --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$e] ensemble', d['value'], d['ms_per_step'])"; done
