"""Sinkhorn scoring time at the bench shape (64 x 2000^2) for the rows-per-work-item given in IMP_SK_ROWS."""
import os, sys, torch
sys.path.insert(0, '.')
from imp_release_b200 import ops, _lib
lib = _lib.load()
B, N = int(os.environ.get('B', 64)), 2000
dist = torch.randn(B, N, N, device='cuda') * 3
bs = torch.tensor(1.0, device='cuda')
ws = ops.SinkhornWorkspace(B, N, N, 'cuda')
for _ in range(3):
    ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
e1.record(); torch.cuda.synchronize()
lib.imp_set_profiling(1)
ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
print(f"rows={os.environ.get('IMP_SK_ROWS', 'auto')} B={B}: scoring {e0.elapsed_time(e1) / 10:.3f} ms, iteration kernel {float(lib.imp_sinkhorn_iter_ms()):.4f} ms", flush=True)
