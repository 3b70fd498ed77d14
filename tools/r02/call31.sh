#!/bin/bash
# round 2, GPU call 31 (8 GPUs): the driver's N = 8 command, default bench (all legs, strong scaling, C5 over 8 GPUs)
mkdir -p gpurun_out/c31
O=gpurun_out/c31
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench 8gpu rc=$?"; tail -c 800 $O/bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c31/bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}); print(d['e2e']); print(d.get('strong_scaling')); print(d.get('e2e_breakdown'))
print({k:(v.get('value'),v.get('e2e_value'),v.get('latency_ms')) for k,v in (d.get('configs') or {}).items()})
PY
