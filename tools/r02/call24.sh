#!/bin/bash
# round 2, GPU call 24: v8 with a suspended (hinted) wait on the expected tile instead of hot polling
mkdir -p gpurun_out/c24
O=gpurun_out/c24
timeout 240 python tools/ab_bench.py pytest v8 tests/test_gpu_kernels.py -m gpu -q -k attention > $O/pytest_v8.log 2>&1; echo "pytest v8 rc=$?"; tail -2 $O/pytest_v8.log
ATTN_TRACE_OUT=tools/_trace/libattn_trace_v8.so timeout 120 python tools/attn_trace.py run > $O/trace_v8.txt 2>&1; echo "trace v8 rc=$?"
sed -n 8,22p $O/trace_v8.txt | cut -c1-110; tail -5 $O/trace_v8.txt
timeout 300 python tools/attn_ab.py base v6r v8 v8p300 v8p4000 > $O/attn_ab.txt 2>&1; echo "attn_ab rc=$?"; tail -6 $O/attn_ab.txt
