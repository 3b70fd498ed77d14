#!/bin/bash
# round 2, GPU call 18: clock64 timeline of attention v8 (one CTA), beside the shipped kernel's
mkdir -p gpurun_out/c18
O=gpurun_out/c18
ATTN_TRACE_OUT=tools/_trace/libattn_trace_v8.so timeout 120 python tools/attn_trace.py run > $O/trace_v8.txt 2>&1; echo "trace v8 rc=$?"
ATTN_TRACE_OUT=tools/_trace/libattn_trace.so timeout 120 python tools/attn_trace.py run > $O/trace_base.txt 2>&1; echo "trace base rc=$?"
cat $O/trace_v8.txt
tail -8 $O/trace_base.txt
