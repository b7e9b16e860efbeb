"""GPU probe: the three storage formats of softmax(M) for the Sinkhorn iteration sweeps (fp32 / fp24 / fp16) on the
bench workload -- isolated scoring time, iteration-kernel time, and the deviation of the model outputs from the fp32
storage on the same inputs."""
import json, os, sys, torch
sys.path.insert(0, '.')
from imp_release_b200 import DGNNS, ops, _lib
from oracle import synth

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
B, N = 64, 2000
ld = (N + 3) // 4 * 4
g = torch.Generator(device='cuda').manual_seed(0)
dist = torch.randn(B, N, ld, device='cuda', generator=g) * 3
bs = torch.tensor(1.0, device='cuda')
lib = _lib.load()
mat = 4.0 * B * (N + 1) * (N + 1)
for fmt in ('legacy', 'fp32', 'fp24', 'fp16'):
    os.environ['IMP_SK_LEGACY'] = '1' if fmt == 'legacy' else '0'
    ws = ops.SinkhornWorkspace(B, N, N, 'cuda', storage='fp32' if fmt == 'legacy' else fmt)
    ms = t(lambda: ops.sinkhorn(dist, ld, bs, 20, ws, write_scores=False), 5)
    lib.imp_set_profiling(1)
    ops.sinkhorn(dist, ld, bs, 20, ws, write_scores=False)
    it_ms = float(lib.imp_sinkhorn_iter_ms())
    lib.imp_set_profiling(0)
    res[f'scoring_{fmt}'] = {'ms': ms, 'iter_kernel_ms': it_ms, 'iter_algorithmic_GBps': mat / it_ms / 1e6}
    print(fmt, res[f'scoring_{fmt}'], flush=True)

nl = 9
cfg = dict(n_layers=nl, GNN_layers=['self', 'cross'] * nl, norm_fn='in', ac_fn='relu', sinkhorn_iterations=20, with_sinkhorn=True, descriptor_dim=256)
sd = synth.make_state_dict('DGNNS', nl, seed=7)
data = {k: (v.cuda() if k.startswith(('desc', 'key', 'sco')) else v) for k, v in synth.make_pair_batch(seed=1, batch=B, n0=N, n1=N).items()}
outs = {}
for fmt in ('legacy', 'fp32', 'fp24', 'fp16'):
    os.environ['IMP_SK_LEGACY'] = '1' if fmt == 'legacy' else '0'
    net = DGNNS({**cfg, 'sinkhorn_storage': 'fp32' if fmt == 'legacy' else fmt}); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    with torch.no_grad():
        o = net(data)
    outs[fmt] = ([x.clone() for x in o['indices0']], [x.clone() for x in o['mscores0']])
    ms = t(lambda: net(data), 3)
    res[f'model_{fmt}_ms_per_step'] = ms
    del net
for fmt in ('fp32', 'fp24', 'fp16'):
    flips = [int((a != b).sum()) for a, b in zip(outs[fmt][0], outs['legacy'][0])]
    dms = [float((a - b).abs().max()) for a, b in zip(outs[fmt][1], outs['legacy'][1])]
    res[f'model_{fmt}_vs_legacy_fp32'] = {'index_flips_per_iteration': flips, 'max_abs_dmscore_per_iteration': dms,
                                   'matches_last': int((outs['legacy'][0][-1] >= 0).sum())}
    print(fmt, res[f'model_{fmt}_vs_legacy_fp32'], flush=True)
print(json.dumps(res))
json.dump(res, open('gpurun_out/sk_formats.json', 'w'), indent=1)
