#!/bin/bash
# round 2, GPU call 14 (1 GPU): evidence for profiles/: launch list of one step, ncu --set full of the top kernels, final full bench
mkdir -p gpurun_out/c14
O=gpurun_out/c14
N="--steps 1 --warmup 1 --no-cpu-baseline --skip-extra-configs"
SOME_B200_BIAS_CORRECTION=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py $N > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
SOME_B200_BIAS_CORRECTION=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair|attention_tc' -s 1 -c 12 -o $O/top python bench.py $N > $O/ncu_top.log 2>&1; echo "ncu top rc=$?"
SOME_B200_BIAS_CORRECTION=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mel_kernel|decode_|dwconv|layernorm' -s 0 -c 8 -o $O/small python bench.py $N > $O/ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench full rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"
ls -la $O
