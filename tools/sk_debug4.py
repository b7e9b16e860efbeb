import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from imp_release_b200 import ops
from oracle import imp_oracle
B, N0, N1, iters = 10, 400, 2500, 5
g = torch.Generator().manual_seed(200 + N0)
dist = torch.randn(B, N0, N1, generator=g) * 3
for b in range(B):
    idx = torch.randperm(min(N0, N1), generator=g)[: min(N0, N1) // 2]
    dist[b, idx, idx] += 12.0
bin_score = torch.tensor(1.3)
ldd = (N1 + 3) // 4 * 4
dd = torch.zeros(B, N0, ldd, device='cuda'); dd[:, :, :N1] = dist.cuda()
ref64 = imp_oracle.sink_algorithm(dist.double(), bin_score.double(), iters)
ref32 = imp_oracle.sink_algorithm(dist, bin_score, iters)
t0, t1, tm0, _ = imp_oracle.compute_matches(ref64.float(), 0.2)
r0, _, rm0, _ = imp_oracle.compute_matches(ref32, 0.2)
print('fp32 oracle vs fp64: flips', int((r0 != t0).sum()))
for fmt in ('fp32', 'fp24'):
    ws = ops.SinkhornWorkspace(B, N0, N1, 'cuda', storage=fmt)
    ops.sinkhorn(dd, ldd, bin_score.cuda(), iters, ws)
    i0, i1, m0, m1 = ops.matches(ws.row_max, ws.row_arg, ws.col_key, 0.2, N0, N1, B)
    bad = (i0.cpu() != t0).nonzero()
    print(fmt, 'flips vs fp64', bad.shape[0], 'max dm', float((m0.cpu() - tm0).abs().max()))
    for b, i in bad.tolist():
        row = ref64[b, i, :-1]
        top = row.topk(2)
        col = ref64[b, :-1, int(top.indices[0])]
        ctop = col.topk(2)
        print('  row', b, i, 'truth', int(t0[b, i]), 'got', int(i0[b, i]), 'row top2', top.values.tolist(), 'col top2', ctop.values.tolist(), ctop.indices.tolist(), 'mscore truth', float(tm0[b, i]), 'got', float(m0[b, i]))
