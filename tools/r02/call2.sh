#!/bin/bash
# round 2, GPU call 2: parity drift dump, attention clock64 trace, ncu --set full (source) of attention / mel / decode / dwconv
mkdir -p gpurun_out/c2
O=gpurun_out/c2
timeout 300 python tools/r02/parity_dump.py $O/parity_dump.npz > $O/parity_dump.log 2>&1; echo "parity_dump rc=$?"; tail -3 $O/parity_dump.log
timeout 120 python tools/attn_trace.py run > $O/attn_trace.txt 2>&1; echo "trace rc=$?"; tail -8 $O/attn_trace.txt
B="--steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 4 -c 1 -o $O/attn python bench.py $B > $O/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mel_kernel|decode|dwconv' -s 3 -c 3 -o $O/small python bench.py $B > $O/ncu_small.log 2>&1; echo "ncu small rc=$?"
ls -la $O
