#!/bin/bash
# round 2, GPU call 25: v8 timeline with the TMA producer's issue times (are the K / V loads late, or issued late?)
mkdir -p gpurun_out/c25
ATTN_TRACE_OUT=tools/_trace/libattn_trace_v8.so timeout 120 python tools/attn_trace.py run > gpurun_out/c25/trace_v8.txt 2>&1; echo "trace v8 rc=$?"
sed -n 22,60p gpurun_out/c25/trace_v8.txt | cut -c1-160
