#!/bin/bash
# Runs every GPU test file in its own process with a timeout (a pipeline bug in a tcgen05 kernel traps after
# ~4 s instead of hanging, but keep the belt and braces) and collects logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/bringup_smi.txt 2>&1
for t in "$@"; do
  name=$(echo "$t" | tr '/:[] ' '_____')
  echo "=== $t" | tee -a gpurun_out/bringup.log
  timeout 300 python -m pytest $t -x -q -m gpu -s --no-header -p no:cacheprovider > gpurun_out/bringup_$name.log 2>&1
  echo "exit=$?" | tee -a gpurun_out/bringup.log
  tail -n 25 gpurun_out/bringup_$name.log | tee -a gpurun_out/bringup.log
done
