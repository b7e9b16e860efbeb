#!/bin/bash
echo "nproc=$(nproc) cpu_count=$(python -c 'import os;print(os.cpu_count())') affinity=$(python -c 'import os;print(len(os.sched_getaffinity(0)))')"
cat /sys/fs/cgroup/cpu.max 2>/dev/null
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" 
for t in 8 16 32 64; do
python - <<PY
import sys, time, torch
sys.path.insert(0,'.')
import bench
torch.set_num_threads($t)
bench.cpu_reference_step(n=1000)
t0=time.perf_counter(); s = bench.cpu_reference_step(n=2000); print("threads", $t, "N=2000 9it: %.2f s/pair" % s, flush=True)
PY
done
