#!/bin/bash
# round 2, GPU call 1: baseline + attention v7 (poly 2/3/4 of 8) + bulk-residual GEMM, all on ONE box (A/B inside one call)
mkdir -p gpurun_out/c1
O=gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > $O/pytest_base.log 2>&1; echo "pytest base rc=$?"; tail -3 $O/pytest_base.log
B="--steps 4 --warmup 3 --no-cpu-baseline"
timeout 600 python bench.py $B > $O/bench_base.json 2> $O/bench_base.err; echo "bench base rc=$?"
for v in v7 v7p3 v7p4; do
  timeout 300 python tools/ab_bench.py pytest $v tests/test_gpu_kernels.py -m gpu -x -q -k attention > $O/pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -2 $O/pytest_$v.log
  timeout 600 python tools/ab_bench.py run $v $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
done
timeout 300 python tools/ab_bench.py pytest bulk tests/test_gpu_kernels.py -m gpu -x -q -k gemm > $O/pytest_bulk.log 2>&1; echo "pytest bulk rc=$?"; tail -2 $O/pytest_bulk.log
timeout 600 python tools/ab_bench.py run bulk $B > $O/bench_bulk.json 2> $O/bench_bulk.err; echo "bench bulk rc=$?"
timeout 600 python bench.py $B > $O/bench_base2.json 2> $O/bench_base2.err; echo "bench base2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c1/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
