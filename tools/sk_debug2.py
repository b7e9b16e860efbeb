import os, sys, torch
sys.path.insert(0, '.')
from imp_release_b200 import ops
torch.manual_seed(0)
B, N0, N1 = 64, 2000, 2000
ldd = 2000
dist = torch.randn(B, N0, ldd, device='cuda') * 3
idx = torch.arange(1000, device='cuda') * 2
dist[:, idx, idx] += 12
bs = torch.tensor(1.3, device='cuda')
os.environ['IMP_SK_LEGACY'] = '0'
ws = ops.SinkhornWorkspace(B, N0, N1, 'cuda', storage='fp32')
ws.q_store.zero_()
us = []
for rep in range(4):
    ops.sinkhorn(dist, ldd, bs, 2, ws, write_scores=False)
    torch.cuda.synchronize()
    us.append(ws.u.clone())
    ldq = 2016
    Q = ws.q_store.view(torch.float32).view(B, N0 + 1, ldq)
    # reference p from iters=0 scores
    if rep == 0:
        w0 = ops.SinkhornWorkspace(B, N0, N1, 'cuda', storage='fp32')
        ops.sinkhorn(dist, ldd, bs, 0, w0, write_scores=True)
        p = w0.scores().clone()
    dq = (Q[:, :, :N1 + 1] - p).abs()
    bad = (dq > 0).nonzero()
    print('rep', rep, 'Q mismatches vs p:', bad.shape[0], 'pad nonzero:', int((Q[:, :, N1 + 1:] != 0).sum()))
    if bad.shape[0]:
        print('  sample', bad[:10].tolist())
        rows = torch.unique(bad[:, :2], dim=0)
        print('  distinct bad rows', rows.shape[0], rows[:10].tolist())
for rep in range(1, 4):
    ru = ((us[rep] - us[0]).abs() / us[0].abs())
    print('u rep', rep, 'vs 0: max rel', float(ru.max()), 'rows >1e-5:', int((ru > 1e-5).sum()))
