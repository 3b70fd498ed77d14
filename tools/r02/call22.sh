#!/bin/bash
# round 2, GPU call 22: interleaved attention A/B (same box, same process); v8 with the published QK tile (phase-aliasing fix)
mkdir -p gpurun_out/c22
O=gpurun_out/c22
timeout 300 python tools/attn_ab.py base v6r v9 v9r > $O/attn_ab.txt 2>&1; echo "attn_ab rc=$?"; cat $O/attn_ab.txt
for v in v8 v8r; do
timeout 240 python tools/ab_bench.py pytest $v tests/test_gpu_kernels.py -m gpu -q -k attention > $O/pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -2 $O/pytest_$v.log
done
ATTN_TRACE_OUT=tools/_trace/libattn_trace_v8.so timeout 120 python tools/attn_trace.py run > $O/trace_v8.txt 2>&1; echo "trace v8 rc=$?"
tail -34 $O/trace_v8.txt
timeout 300 python tools/attn_ab.py base v9 v8 v8r > $O/attn_ab2.txt 2>&1; echo "attn_ab2 rc=$?"; tail -6 $O/attn_ab2.txt
