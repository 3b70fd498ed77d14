#!/bin/bash
# round 2, GPU call 21: interleaved attention A/B (same box, same process), v8 after the s_full phase fix, GEMM role remap
mkdir -p gpurun_out/c21
O=gpurun_out/c21
for v in v8 v8r; do
timeout 240 python tools/ab_bench.py pytest $v tests/test_gpu_kernels.py -m gpu -q -k attention > $O/pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -2 $O/pytest_$v.log
done
ATTN_TRACE_OUT=tools/_trace/libattn_trace_v8.so timeout 120 python tools/attn_trace.py run > $O/trace_v8.txt 2>&1; echo "trace v8 rc=$?"
tail -30 $O/trace_v8.txt
timeout 300 python tools/attn_ab.py base v6r v9 v9r v8 v8r > $O/attn_ab.txt 2>&1; echo "attn_ab rc=$?"; cat $O/attn_ab.txt
timeout 240 python tools/ab_bench.py pytest gemmr tests/test_gpu_kernels.py -m gpu -q -k gemm > $O/pytest_gemmr.log 2>&1; echo "pytest gemmr rc=$?"; tail -2 $O/pytest_gemmr.log
B="--steps 3 --warmup 3 --no-cpu-baseline --skip-extra-configs"
timeout 300 python bench.py $B > $O/bench_base.json 2> $O/bench_base.err; echo "bench base rc=$?"
timeout 300 python tools/ab_bench.py run gemmr $B > $O/bench_gemmr.json 2> $O/bench_gemmr.err; echo "bench gemmr rc=$?"
timeout 300 python bench.py $B > $O/bench_base2.json 2> $O/bench_base2.err; echo "bench base rc=$?"
timeout 300 python tools/ab_bench.py run gemmr $B > $O/bench_gemmr2.json 2> $O/bench_gemmr2.err; echo "bench gemmr rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c21/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']; pc=d.get('parity_check') or {}
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'], 'parity', pc.get('max_abs_probs'), pc.get('note_frame_agreement'))
        print('   ', ' '.join('%s=%.3f'%(n,v['ms_per_step']) for n,v in d['gemm_shapes'].items()))
    except Exception as e:
        print(f, 'ERR', e)
PY
