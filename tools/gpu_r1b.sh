#!/bin/bash
# GPU session: compact-storage Sinkhorn -- tests, format probe, bench per format.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "sinkhorn" -x > gpurun_out/t_sinkhorn.log 2>&1
echo "sinkhorn tests rc=$?"; tail -n 15 gpurun_out/t_sinkhorn.log
timeout 600 python tools/sk_formats.py > gpurun_out/sk_formats.log 2>&1
echo "sk_formats rc=$?"; tail -n 12 gpurun_out/sk_formats.log
for f in fp32 fp24 fp16; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sinkhorn-storage $f > gpurun_out/bench_$f.json 2> gpurun_out/bench_$f.err
  echo "bench $f rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['kernel'][:30], round(d['roofline']['frac'],3), {k:v['ms_per_step_share'] for k,v in list(d['kernels'].items())[:6]})
" 2>&1 | tail -3
done
