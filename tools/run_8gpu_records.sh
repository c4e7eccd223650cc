TAG=${1:-r02e}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
S="--gpus 8 --steps 20 --warmup 5 --no-workloads --no-cpu-baseline --no-other-modes"
timeout 300 $TR --master-port 29521 bench.py $S > gpurun_out/${TAG}_bench_8gpu_weak.json 2> gpurun_out/${TAG}_8gpu_weak.err
timeout 300 $TR --master-port 29522 bench.py $S --scaling strong --sustain-s 0 > gpurun_out/${TAG}_bench_8gpu_strong.json 2> gpurun_out/${TAG}_8gpu_strong.err
timeout 300 $TR --master-port 29523 bench.py $S --workload bridge_encode > gpurun_out/${TAG}_bench_8gpu_encode.json 2> gpurun_out/${TAG}_8gpu_encode.err
timeout 300 $TR --master-port 29524 bench.py $S --workload ensemble > gpurun_out/${TAG}_bench_8gpu_ensemble.json 2> gpurun_out/${TAG}_8gpu_ensemble.err
for f in weak strong encode ensemble; do python -c "
import json,sys; d=json.loads(open('gpurun_out/${TAG}_bench_8gpu_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), (d.get('rank_check') or {}).get('ok'), d.get('roofline',{}).get('frac'))"; done
