"""Secondary BASELINE.json configs on one B200 (bench.py covers configs[1]): EIMP batched pruning (configs[2]),
B=1 YFCC-shape latency through produce_matches(only_last=True) and the per-layer API (configs[3], one rank),
Sinkhorn-only microbench (configs[4]).  Prints one JSON object; numbers are CUDA-event timed, 3 warm-ups."""
import json, sys, time
import torch
sys.path.insert(0, '.')
from imp_release_b200 import DGNNS, AdaGMN, ops, normalize_keypoints
from oracle import synth

def cfg(nl): return dict(n_layers=nl, GNN_layers=['self', 'cross'] * nl, norm_fn='in', ac_fn='relu', sinkhorn_iterations=20, with_sinkhorn=True, n_min_tokens=256)
def timed(fn, warm=3, n=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {}
dev = 'cuda'
# ---- configs[2]: EIMP, N=2000 -> pruned, 9 iters, batch 128
B = 128
net = AdaGMN(cfg(9)); net.load_state_dict(synth.make_state_dict('AdaGMN', 9, seed=7, bin_score=8.0)); net = net.to(dev).eval()
data = {k: v.to(dev) for k, v in synth.make_pair_batch(seed=2, batch=B, n0=2000, n1=2000).items()}
with torch.no_grad():
    ms = timed(lambda: net(data), warm=2, n=3)
    cnt, _ = net._kept
out['eimp_b128_n2000_9it'] = {'ms_per_batch': ms, 'pairs_per_s': B / ms * 1e3, 'kept_mean': float(cnt.float().mean()), 'kept_min': int(cnt.min()), 'kept_max': int(cnt.max()), 'bin_score': 8.0}
del net, data; torch.cuda.empty_cache()
# ---- configs[3] (one rank): DGNNS 15 iters, B=1, ragged N, produce_matches(only_last=True)
net = DGNNS(cfg(15)); net.load_state_dict(synth.make_state_dict('DGNNS', 15, seed=7)); net = net.to(dev).eval()
data = {k: v.to(dev) for k, v in synth.make_pair_batch(seed=5, batch=1, n0=1987, n1=1733, width=1600, height=1200).items()}
with torch.no_grad():
    ms = timed(lambda: net.produce_matches(data, p=0.2, only_last=True), warm=3, n=10)
out['imp_b1_15it_only_last'] = {'ms_per_pair': ms, 'pairs_per_s': 1e3 / ms}
from imp_release_b200.graphed import GraphedMatcher
g = GraphedMatcher(net, data, p=0.2, only_last=True)
ms = timed(lambda: g(data), warm=3, n=20)
out['imp_b1_15it_only_last_cuda_graph'] = {'ms_per_pair': ms, 'pairs_per_s': 1e3 / ms}
# per-layer API (eval/matching.py sequence without the host RANSAC): encode + 15 x (self, cross) + 7 scorings
def layer_api():
    nk0 = normalize_keypoints(data['keypoints0'], data['image0'].shape); nk1 = normalize_keypoints(data['keypoints1'], data['image1'].shape)
    e0, e1 = net.encode_keypoint(nk0, nk1, data['scores0'], data['scores1'])
    d0 = data['descriptors0'].transpose(1, 2) + e0; d1 = data['descriptors1'].transpose(1, 2) + e1
    for it in range(15):
        d0, d1 = net.forward_one_layer(d0, d1, None, None, 2 * it)
        d0, d1 = net.forward_one_layer(d0, d1, None, None, 2 * it + 1)
        if it in (3, 5, 7, 9, 11, 13, 14):
            dist = net.compute_distance(d0, d1, layer_id=it)
            sc = net.compute_score(dist, net.bin_score, net.sinkhorn_iterations)
            i0, i1, m0, m1 = net.compute_matches(sc, p=0.1)
    return i0
with torch.no_grad():
    ms = timed(layer_api, warm=2, n=5)
out['imp_b1_15it_layer_api_7_scorings'] = {'ms_per_pair': ms, 'pairs_per_s': 1e3 / ms}
del net; torch.cuda.empty_cache()
# ---- configs[4]: Sinkhorn only, 2048^2 (dist 2047^2), 100 iterations
for Bs, storage in ((1, None), (16, 'fp32'), (16, 'fp24'), (16, 'fp16')):
    N = 2047; ld = 2048
    dist = torch.randn(Bs, N, ld, device=dev) * 3
    ws = ops.SinkhornWorkspace(Bs, N, N, dev, storage=storage); bs = torch.tensor(1.0, device=dev)
    ms = timed(lambda: ops.sinkhorn(dist, ld, bs, 100, ws, write_scores=True), warm=2, n=3)
    mat = 4.0 * Bs * 2048 * 2048
    peak = 6543.7   # MEASURED_PEAKS.json hbm_gbs
    out[f'sinkhorn_2048sq_100it_b{Bs}' + (f'_{storage}' if storage else '')] = {
        'ms': ms, 'algorithmic_GBps': 2 * 100 * mat / ms / 1e6, 'sweeps_GBps_at_4B_per_element': (100 + 2) * mat / ms / 1e6,
        'frac_of_measured_hbm': 2 * 100 * mat / ms / 1e6 / peak,
        'path': 'shared-memory resident (cooperative)' if ws.q_store is None else f'column-split streaming, {storage} copy'}
print(json.dumps(out, indent=1))
