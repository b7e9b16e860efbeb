#!/bin/bash
# GPU session: full GPU test-suite, default bench, ncu evidence (tag r01f)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$?"; tail -n 6 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel'][:40], round(d['roofline']['frac'],3), d['clocks'])
print({k:v['ms_per_step_share'] for k,v in list(d['kernels'].items())[:8]})
print(d['cpu_baseline'])
PY
bash tools/ncu_capture.sh r01f 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sinkhorn-storage fp32 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fp32.json').read().strip().splitlines()[-1])
print('fp32 storage:', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
