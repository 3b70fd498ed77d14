#!/bin/bash
# round 2, GPU call 28 (4 GPUs): e2e after the host-side changes of the sharded path (memoised layouts, no copy of the gathered
# slab, conversions restricted to the used rows)
mkdir -p gpurun_out/c28
O=gpurun_out/c28
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 5 --warmup 3 --skip-extra-configs > $O/bench_4gpu.json 2> $O/bench_4gpu.err; echo "bench 4gpu rc=$?"; tail -c 300 $O/bench_4gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c28/bench_4gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}); print(d['e2e']); print(d.get('strong_scaling')); print(d.get('e2e_breakdown')); print(d.get('parity_check'))
PY
