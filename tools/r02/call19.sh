#!/bin/bash
# round 2, GPU call 19: attention v8 with a non-blocking MMA-thread poll (test_wait + hinted try_wait); timeline + A/B
mkdir -p gpurun_out/c19
O=gpurun_out/c19
ATTN_TRACE_OUT=tools/_trace/libattn_trace_v8.so timeout 120 python tools/attn_trace.py run > $O/trace_v8.txt 2>&1; echo "trace v8 rc=$?"
tail -8 $O/trace_v8.txt
timeout 240 python tools/ab_bench.py pytest v8 tests/test_gpu_kernels.py -m gpu -q -k attention > $O/pytest_v8.log 2>&1; echo "pytest v8 rc=$?"; tail -3 $O/pytest_v8.log
B="--steps 3 --warmup 3 --no-cpu-baseline --skip-extra-configs"
timeout 300 python bench.py $B > $O/bench_base.json 2> $O/bench_base.err; echo "bench base rc=$?"
for v in v8 v8p20 v8p1000; do
timeout 300 python tools/ab_bench.py run $v $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
done
timeout 300 python bench.py $B > $O/bench_base2.json 2> $O/bench_base2.err; echo "bench base rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c19/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']; pc=d.get('parity_check') or {}
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'], 'parity', pc.get('max_abs_probs'), pc.get('note_frame_agreement'))
    except Exception as e:
        print(f, 'ERR', e)
PY
