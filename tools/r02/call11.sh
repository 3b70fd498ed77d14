#!/bin/bash
# round 2, GPU call 11: attention clock64 timeline (normal + hand-off chain alone), tests
mkdir -p gpurun_out/c11
O=gpurun_out/c11
timeout 120 python tools/attn_trace.py run > $O/attn_trace.txt 2>&1; echo "trace rc=$?"; tail -8 $O/attn_trace.txt
ATTN_TRACE_OUT=tools/_trace/libattn_trace_nosm.so timeout 120 python tools/attn_trace.py run > $O/attn_trace_nosoftmax.txt 2>&1; echo "trace nosoftmax rc=$?"; tail -8 $O/attn_trace_nosoftmax.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/pytest_all.log
