#!/bin/bash
# round 2, GPU call 15: attention v8 (shared O accumulator, three S buffers, rendezvous for the reference maximum)
mkdir -p gpurun_out/c15
O=gpurun_out/c15
timeout 240 python tools/ab_bench.py pytest v8 tests/test_gpu_kernels.py -m gpu -x -q -k attention > $O/pytest_v8.log 2>&1; echo "pytest v8 rc=$?"; tail -5 $O/pytest_v8.log
B="--steps 3 --warmup 3 --no-cpu-baseline --skip-extra-configs"
timeout 300 python bench.py $B > $O/bench_base.json 2> $O/bench_base.err; echo "bench base rc=$?"
timeout 300 python tools/ab_bench.py run v8 $B > $O/bench_v8.json 2> $O/bench_v8.err; echo "bench v8 rc=$?"; tail -3 $O/bench_v8.err
timeout 300 python tools/ab_bench.py pytest v8 tests/test_gpu_pipeline.py tests/test_gpu_parity_long.py -m gpu -q > $O/pytest_v8_pipe.log 2>&1; echo "pytest v8 pipeline rc=$?"; tail -3 $O/pytest_v8_pipe.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c15/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']; pc=d.get('parity_check') or {}
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'], 'parity', pc.get('max_abs_probs'), pc.get('mean_bounds_error'), pc.get('note_frame_agreement'), pc.get('note_exact_boundary_agreement'))
    except Exception as e:
        print(f, 'ERR', e)
PY
