import sys, torch
sys.path.insert(0, '.')
from imp_release_b200 import ops
fmt = sys.argv[1] if len(sys.argv) > 1 else 'fp24'
B, N = 64, 2000
dist = torch.randn(B, N, N, device='cuda') * 3
bs = torch.tensor(1.0, device='cuda')
ws = ops.SinkhornWorkspace(B, N, N, 'cuda', storage=fmt)
for _ in range(2):
    ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
torch.cuda.synchronize()
