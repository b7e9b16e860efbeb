#!/bin/bash
# Runs on the B200 box (via gpurun): each test group in its own process with a hard timeout, so a hung kernel
# cannot take the rest of the session down.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
for t in "$@"; do
  name=$(echo "$t" | tr '/:[], ' '______')
  echo "=== $t" | tee -a gpurun_out/session.log
  timeout 600 python -m pytest "$t" -m gpu -q --no-header -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "rc=$rc" | tee -a gpurun_out/session.log
  tail -n 25 "gpurun_out/$name.log" | tee -a gpurun_out/session.log
done
