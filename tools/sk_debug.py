import os, sys, torch
sys.path.insert(0, '.')
from imp_release_b200 import ops
sys.path.insert(0, 'tests')
torch.manual_seed(0)
B, N0, N1 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fmt = sys.argv[4]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20
ldd = (N1 + 3) // 4 * 4
dist = torch.randn(B, N0, ldd, device='cuda') * 3
idx = torch.arange(min(N0, N1) // 2, device='cuda') * 2
dist[:, idx, idx] += 12
bs = torch.tensor(1.3, device='cuda')
os.environ['IMP_SK_LEGACY'] = '1'
wl = ops.SinkhornWorkspace(B, N0, N1, 'cuda', storage='fp32')
ops.sinkhorn(dist, ldd, bs, iters, wl, write_scores=True)
ref = wl.scores().clone()
os.environ['IMP_SK_LEGACY'] = '0'
ws = ops.SinkhornWorkspace(B, N0, N1, 'cuda', storage=fmt)
assert ws.q_store is not None
ops.sinkhorn(dist, ldd, bs, iters, ws, write_scores=True)
sc = ws.scores()
d = (sc - ref).abs()
print('max abs diff', float(d.max()), 'u diff', float((ws.u - wl.u).abs().max() / wl.u.abs().max()))
ru = ((ws.u - wl.u).abs() / wl.u.abs())
bad = (ru > 1e-3).nonzero()
print('rows with bad u:', bad.shape[0], 'of', ru.numel())
if bad.shape[0]:
    print('first bad (b, i):', bad[:20].tolist())
    rows = bad[:, 1]
    print('row index histogram mod 8:', torch.bincount(rows % 8, minlength=8).tolist())
    print('batches affected:', torch.unique(bad[:, 0]).tolist()[:40])
    print('min/max row', int(rows.min()), int(rows.max()))
print('row_arg mismatch', int((ws.row_arg != wl.row_arg).sum()), 'col_key mismatch', int(((ws.col_key & 0xFFFFFFFF) != (wl.col_key & 0xFFFFFFFF)).sum()))
