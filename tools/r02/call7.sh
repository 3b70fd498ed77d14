#!/bin/bash
# round 2, GPU call 7: LN-fold cost diagnosis (timing-only variants), ncu of the new mel + consumer / producer GEMMs, full new bench.py
mkdir -p gpurun_out/c7
O=gpurun_out/c7
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention or mel or decode" > $O/pytest_k.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_k.log
B="--steps 3 --warmup 3 --no-cpu-baseline --skip-extra-configs"
timeout 600 python bench.py $B > $O/bench_fold.json 2> $O/bench_fold.err; echo "bench fold rc=$?"
for v in lnc_nostats lnc_nos lnp_noxb lnp_nostats; do
  timeout 600 python tools/ab_bench.py run $v $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c7/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'], d.get('parity_check'))
        print('    ', ' '.join('%s=%.3f'%(n.replace('_N','/').replace('_K','/'),v['ms_per_step']) for n,v in sorted(d.get('gemm_shapes',{}).items())))
    except Exception as e:
        print(f, 'ERR', e)
PY
N="--steps 1 --warmup 1 --no-cpu-baseline --skip-extra-configs"
SOME_B200_BIAS_CORRECTION=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 0 -c 6 -o $O/gemm python bench.py $N > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
SOME_B200_BIAS_CORRECTION=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 1 -c 1 -o $O/mel python bench.py $N > $O/ncu_mel.log 2>&1; echo "ncu mel rc=$?"
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench full rc=$?"; tail -c 1500 $O/bench_full.err
ls -la $O | head -30
