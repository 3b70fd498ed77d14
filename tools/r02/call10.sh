#!/bin/bash
# round 2, GPU call 10: pipelined + TMA-stored GEMM epilogues, head-level calibration, attention decomposition (timing-only builds)
mkdir -p gpurun_out/c10
O=gpurun_out/c10
timeout 900 python -m pytest tests -m gpu -q -s > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; grep -E "agreement|passed|failed|FAILED" $O/pytest_all.log | tail -14
B="--steps 3 --warmup 3 --no-cpu-baseline --skip-extra-configs"
timeout 600 python bench.py $B > $O/bench_base.json 2> $O/bench_base.err; echo "bench base rc=$?"
for v in att_nomma att_nosoftmax; do
  timeout 600 python tools/ab_bench.py run $v $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
done
timeout 600 python bench.py $B > $O/bench_base2.json 2> $O/bench_base2.err; echo "bench base2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c10/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']; pc=d.get('parity_check') or {}
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'], 'parity', pc.get('max_abs_probs'), pc.get('mean_bounds_error'), pc.get('note_frame_agreement'), pc.get('note_exact_boundary_agreement'))
        print('    ', ' '.join('%s=%.3f/%.2f'%(n.replace('_N','/').replace('_K','/'),v['ms_per_step'],v['frac']) for n,v in sorted(d.get('gemm_shapes',{}).items())))
    except Exception as e:
        print(f, 'ERR', e)
PY
