#!/bin/bash
# ncu --set full of one launch of each attention variant at the bench shape (tools/attn_probe.py)
for v in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:attention" -s 4 -c 1 -f -o gpurun_out/attn_var${v} python tools/attn_probe.py $v > gpurun_out/attn_var${v}.log 2>&1
  tail -2 gpurun_out/attn_var${v}.log
done
