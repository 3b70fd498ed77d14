#!/bin/bash
# round 2, GPU call 9: defaults (unfolded LN, mel 3 CTAs), key-shift mel, stale-max attention A/B, small-batch device timing, sanitizer
mkdir -p gpurun_out/c9
O=gpurun_out/c9
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 $O/pytest_all.log
B="--steps 3 --warmup 3 --no-cpu-baseline --skip-extra-configs"
timeout 600 python bench.py $B > $O/bench_base.json 2> $O/bench_base.err; echo "bench base rc=$?"
timeout 300 python tools/ab_bench.py pytest stale tests/test_gpu_kernels.py -m gpu -x -q -k attention > $O/pytest_stale.log 2>&1; echo "pytest stale rc=$?"; tail -2 $O/pytest_stale.log
timeout 600 python tools/ab_bench.py run stale $B > $O/bench_stale.json 2> $O/bench_stale.err; echo "bench stale rc=$?"
timeout 600 python bench.py $B > $O/bench_base2.json 2> $O/bench_base2.err; echo "bench base2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c9/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']; pc=d.get('parity_check') or {}
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'], 'parity', pc.get('max_abs_probs'), pc.get('mean_bounds_error'), pc.get('note_frame_agreement'), pc.get('note_exact_boundary_agreement'))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 python tools/r02/small_batch.py > $O/small_batch.txt 2>&1; echo "small batch rc=$?"; grep -E "graphs (True|False)" $O/small_batch.txt
# compute-sanitizer over the kernel tests (memcheck, then racecheck on the shared-memory heavy kernels)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or attention or mel or decode or dwconv or layernorm or row_stats" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Error" $O/sanitizer_memcheck.log | tail -5
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mel or decode or dwconv or layernorm or row_stats or gemm_resid or ln_producer" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|Error" $O/sanitizer_racecheck.log | tail -5
