#!/bin/bash
# round 2, GPU call 8: cache-hinted consumer epilogues (+ 4-stage variant), mel v2b, fp32 validation mode, CUDA graphs
mkdir -p gpurun_out/c8
O=gpurun_out/c8
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_accurate.py -m gpu -x -q -s > $O/pytest_kpa.log 2>&1; echo "pytest rc=$?"; grep -E "max\|probs|passed|failed|Error" $O/pytest_kpa.log | tail -12
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -4 $O/pytest_all.log
B="--steps 3 --warmup 3 --no-cpu-baseline --skip-extra-configs"
timeout 600 python bench.py $B > $O/bench_fold.json 2> $O/bench_fold.err; echo "bench fold rc=$?"
timeout 600 python tools/ab_bench.py run lnc4 $B > $O/bench_lnc4.json 2> $O/bench_lnc4.err; echo "bench lnc4 rc=$?"
timeout 600 python tools/ab_bench.py run mel3 $B > $O/bench_mel3.json 2> $O/bench_mel3.err; echo "bench mel3 rc=$?"
SOME_B200_LN_FOLD=0 timeout 600 python bench.py $B > $O/bench_nofold.json 2> $O/bench_nofold.err; echo "bench nofold rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c8/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']; pc=d.get('parity_check') or {}
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'], 'parity', pc.get('mean_bounds_error'), pc.get('note_frame_agreement'), pc.get('note_exact_boundary_agreement'))
        print('    ', ' '.join('%s=%.3f'%(n.replace('_N','/').replace('_K','/'),v['ms_per_step']) for n,v in sorted(d.get('gemm_shapes',{}).items())))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 python tools/r02/small_batch.py > $O/small_batch.txt 2>&1; echo "small batch rc=$?"; cat $O/small_batch.txt | tail -12
