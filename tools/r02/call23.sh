#!/bin/bash
# round 2, GPU call 23: v8 with the operand waits off the hand-off chain + all pending P tiles rescaled; product = no-max-pass kernel
mkdir -p gpurun_out/c23
O=gpurun_out/c23
timeout 240 python tools/ab_bench.py pytest v8 tests/test_gpu_kernels.py -m gpu -q -k attention > $O/pytest_v8.log 2>&1; echo "pytest v8 rc=$?"; tail -2 $O/pytest_v8.log
ATTN_TRACE_OUT=tools/_trace/libattn_trace_v8.so timeout 120 python tools/attn_trace.py run > $O/trace_v8.txt 2>&1; echo "trace v8 rc=$?"
tail -12 $O/trace_v8.txt
timeout 300 python tools/attn_ab.py base v6r v8 > $O/attn_ab.txt 2>&1; echo "attn_ab rc=$?"; tail -6 $O/attn_ab.txt
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_product.log 2>&1; echo "pytest product rc=$?"; tail -3 $O/pytest_product.log
