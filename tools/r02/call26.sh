#!/bin/bash
# round 2, GPU call 26: programmatic dependent launch of the trunk kernels: correctness (whole GPU suite with PDL forced on),
# small-batch latency (graphs x LayerNorm folding x PDL), effect on the large batch
mkdir -p gpurun_out/c26
O=gpurun_out/c26
SOME_B200_PDL=all timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_pdl_all.log 2>&1; echo "pytest PDL=all rc=$?"; tail -3 $O/pytest_pdl_all.log
timeout 600 python tools/r02/small_batch.py > $O/small_batch.txt 2>&1; echo "small_batch rc=$?"; grep -v "^{" $O/small_batch.txt | tail -20
B="--steps 3 --warmup 3 --no-cpu-baseline --skip-extra-configs"
SOME_B200_PDL=off timeout 300 python bench.py $B > $O/bench_pdl_off.json 2> $O/bench_pdl_off.err; echo "bench off rc=$?"
SOME_B200_PDL=all timeout 300 python bench.py $B > $O/bench_pdl_all.json 2> $O/bench_pdl_all.err; echo "bench all rc=$?"
SOME_B200_PDL=off timeout 300 python bench.py $B > $O/bench_pdl_off2.json 2> $O/bench_pdl_off2.err; echo "bench off rc=$?"
SOME_B200_PDL=all timeout 300 python bench.py $B > $O/bench_pdl_all2.json 2> $O/bench_pdl_all2.err; echo "bench all rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c26/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']; pc=d.get('parity_check') or {}
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], ' '.join('%s=%.2f'%(n.replace('some_',''),v['ms_per_step']) for n,v in k.items()), d['clocks']['sm_mhz'], 'parity', pc.get('max_abs_probs'), pc.get('note_frame_agreement'))
    except Exception as e:
        print(f, 'ERR', e)
PY
