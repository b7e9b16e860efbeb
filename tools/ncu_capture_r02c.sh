#!/bin/bash
# Round-2 final evidence: ncu --set full of the top kernels inside the bench workload, the SuperPoint kernels, and a launch list.
tag=${1:-r02c}
mkdir -p gpurun_out
for spec in "attention_kernel:3" "skq_iter_kernel:6" "gemm_f16split:40"; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 500 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c 2 -f -o gpurun_out/${tag}_$k python bench.py --ncu --warmup 1 > gpurun_out/${tag}_ncu_$k.log 2>&1
  tail -1 gpurun_out/${tag}_ncu_$k.log
done
for spec in "conv3x3_kernel:0" "conv1a_kernel:0"; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c 2 -f -o gpurun_out/${tag}_$k python tools/sp_bench.py 1200 1600 > gpurun_out/${tag}_ncu_$k.log 2>&1
  tail -1 gpurun_out/${tag}_ncu_$k.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --ncu --warmup 1 > gpurun_out/${tag}_ncu_launches.log 2>&1
wc -l gpurun_out/${tag}_launches.csv
