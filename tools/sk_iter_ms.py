import sys, torch
sys.path.insert(0, '.')
from imp_release_b200 import ops, _lib
fmt = sys.argv[1] if len(sys.argv) > 1 else 'fp24'
B, N = 64, 2000
dist = torch.randn(B, N, N, device='cuda') * 3
bs = torch.tensor(1.0, device='cuda')
ws = ops.SinkhornWorkspace(B, N, N, 'cuda', storage=fmt)
lib = _lib.load()
for _ in range(3):
    ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
e1.record(); torch.cuda.synchronize()
lib.imp_set_profiling(1)
ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
it = float(lib.imp_sinkhorn_iter_ms())
print(f'{fmt}: scoring {e0.elapsed_time(e1) / 5:.3f} ms, iteration kernel {it:.4f} ms')
