#!/bin/bash
# round 2, GPU call 5: L1-prefetched LN stats, bias correction (parity-long report), attention v6b A/B
mkdir -p gpurun_out/c5
O=gpurun_out/c5
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest_kp.log 2>&1; echo "pytest kernels+pipeline rc=$?"; tail -3 $O/pytest_kp.log
timeout 900 python -m pytest tests/test_gpu_parity_long.py -m gpu -q -s > $O/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; grep -E "agreement|passed|failed" $O/pytest_parity.log
SOME_B200_BIAS_CORRECTION=0 timeout 900 python -m pytest tests/test_gpu_parity_long.py -m gpu -q -s > $O/pytest_parity_nobc.log 2>&1; echo "pytest parity (no bias correction) rc=$?"; grep -E "agreement|passed|failed" $O/pytest_parity_nobc.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -6 $O/pytest_all.log
B="--steps 4 --warmup 3 --no-cpu-baseline"
timeout 600 python bench.py $B > $O/bench_fold.json 2> $O/bench_fold.err; echo "bench fold rc=$?"
SOME_B200_LN_FOLD=0 timeout 600 python bench.py $B > $O/bench_nofold.json 2> $O/bench_nofold.err; echo "bench nofold rc=$?"
timeout 300 python tools/ab_bench.py pytest v6b tests/test_gpu_kernels.py -m gpu -x -q -k attention > $O/pytest_v6b.log 2>&1; echo "pytest v6b rc=$?"; tail -2 $O/pytest_v6b.log
timeout 600 python tools/ab_bench.py run v6b $B > $O/bench_v6b.json 2> $O/bench_v6b.err; echo "bench v6b rc=$?"
timeout 600 python bench.py $B > $O/bench_fold2.json 2> $O/bench_fold2.err; echo "bench fold2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c5/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'])
        for n,v in sorted(d.get('gemm_shapes',{}).items()): print('    ', n, v['ms_per_step'], v['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
