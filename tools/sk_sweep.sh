#!/bin/bash
# Sinkhorn micro-sweep on the GPU box: config 5 of BASELINE.json (2048^2, 100 iters) and the bench shape, across L2 budgets
mkdir -p gpurun_out
for mb in 0 48 80 100; do
  IMP_SK_L2_MB=$mb python - <<PY
import os, sys, time, torch
sys.path.insert(0, '.')
from imp_release_b200 import ops
def run(B, N, iters, write):
    dist = torch.randn(B, N, N, device='cuda') * 3
    ws = ops.SinkhornWorkspace(B, N, N, 'cuda')
    bs = torch.tensor(1.0, device='cuda')
    for _ in range(2): ops.sinkhorn(dist, N, bs, iters, ws, write_scores=write)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ops.sinkhorn(dist, N, bs, iters, ws, write_scores=write)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    mat = 4.0 * B * (N + 1) * (N + 1)
    print(f"L2_MB={os.environ['IMP_SK_L2_MB']:>4s} B={B:3d} N={N} iters={iters:3d} write={int(write)}: {ms:8.3f} ms  "
          f"algorithmic {(2*iters*mat)/ms/1e6:8.1f} GB/s  actual-sweeps {((iters+3)*mat)/ms/1e6:8.1f} GB/s", flush=True)
run(64, 2000, 20, False)
run(16, 2047, 100, True)
run(1, 2047, 100, True)
PY
done
