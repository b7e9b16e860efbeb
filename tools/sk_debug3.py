import os, sys, torch
sys.path.insert(0, '.')
from imp_release_b200 import ops
torch.manual_seed(0)
B, N0, N1 = int(sys.argv[1]), 2000, 2000
fmt = sys.argv[2]; reps = int(sys.argv[3])
ldd = 2000
dist = torch.randn(B, N0, ldd, device='cuda') * 3
idx = torch.arange(1000, device='cuda') * 2
dist[:, idx, idx] += 12
bs = torch.tensor(1.3, device='cuda')
os.environ['IMP_SK_LEGACY'] = '0'
ws = ops.SinkhornWorkspace(B, N0, N1, 'cuda', storage=fmt)
us = []
for rep in range(reps):
    ops.sinkhorn(dist, ldd, bs, 2, ws, write_scores=False)
    torch.cuda.synchronize()
    us.append(ws.u.clone())
for rep in range(1, reps):
    ru = ((us[rep] - us[0]).abs() / us[0].abs())
    print(fmt, 'u rep', rep, 'vs 0: max rel', float(ru.max()), 'rows >1e-5:', int((ru > 1e-5).sum()))
