#!/bin/bash
# round 2, GPU call 30 (1 GPU): final evidence of the round-2 build: GPU test-suite, smoke(), launch list of one step, ncu --set full of
# the attention kernel and the GEMM variants, the default bench line, the reference arm
mkdir -p gpurun_out/c30
O=gpurun_out/c30
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -2 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
N="--steps 1 --warmup 1 --no-cpu-baseline --skip-extra-configs"
SOME_B200_BIAS_CORRECTION=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py $N > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
SOME_B200_BIAS_CORRECTION=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair|attention_tc' -s 1 -c 12 -o $O/top python bench.py $N > $O/ncu_top.log 2>&1; echo "ncu top rc=$?"
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench full rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"
tail -c 1500 $O/bench_full.json
ls -la $O
