#!/bin/bash
# round 2, GPU call 12 (2 GPUs): full default bench under torchrun (dist changes: in-place gather slot, pinned landing, lazy unpack; C3/C4/C5 legs at N=2)
mkdir -p gpurun_out/c12
O=gpurun_out/c12
nvidia-smi -L > $O/gpus.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -c 600 $O/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c12/bench_2gpu.json').read().strip().splitlines()[-1])
print('value %.0f  e2e %.0f  ms %.2f  e2e_ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))
print('breakdown', d.get('e2e_breakdown'))
print('strong', d.get('strong_scaling'))
for k,v in (d.get('configs') or {}).items(): print(k, {a:b for a,b in v.items() if a in ('value','e2e_value','latency_ms','chunks','error','ms_per_step','e2e_ms_per_step')})
PY
