#!/bin/bash
# round 2, GPU call 29: share of exps on the FMA-pipe polynomial, re-tuned for the no-max-pass kernel (0 / 2 / 4 of every 8)
mkdir -p gpurun_out/c29
ATTN_AB_ROUNDS=16 timeout 300 python tools/attn_ab.py base poly0 poly4 > gpurun_out/c29/attn_ab.txt 2>&1; echo "attn_ab rc=$?"; cat gpurun_out/c29/attn_ab.txt
