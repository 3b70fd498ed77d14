#!/bin/bash
# round 2, GPU call 4: TMA residual epilogue + LayerNorm folding — kernel tests, pipeline tests, bench fold on / off
mkdir -p gpurun_out/c4
O=gpurun_out/c4
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or row_stats or layernorm" > $O/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -5 $O/pytest_gemm.log
timeout 900 python -m pytest tests -m gpu -q -s > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 $O/pytest_all.log
B="--steps 4 --warmup 3 --no-cpu-baseline"
timeout 600 python bench.py $B > $O/bench_fold.json 2> $O/bench_fold.err; echo "bench fold rc=$?"
SOME_B200_LN_FOLD=0 timeout 600 python bench.py $B > $O/bench_nofold.json 2> $O/bench_nofold.err; echo "bench nofold rc=$?"
timeout 600 python bench.py $B > $O/bench_fold2.json 2> $O/bench_fold2.err; echo "bench fold2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c4/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'])
        for n,v in sorted(d.get('gemm_shapes',{}).items()): print('    ', n, v)
    except Exception as e:
        print(f, 'ERR', e)
PY
