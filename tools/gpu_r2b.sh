#!/bin/bash
# GPU session (round 2, second half): full GPU test-suite, smoke, default bench (with secondary configs incl. SuperPoint), reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$?"; tail -n 6 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02f.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel'][:40], round(d['roofline']['frac'],3), d['clocks'])
print({k: round(v['frac'],3) for k,v in d['roofline_other'].items()})
print(d['cpu_baseline'])
print(json.dumps({k: v for k, v in d['secondary'].items() if 'EIMP' in k or 'image pair' in k}, indent=1)[:1500])
PY
tail -3 gpurun_out/bench_r02f.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02f_ref.json 2>/dev/null; cat gpurun_out/bench_r02f_ref.json | cut -c1-300
