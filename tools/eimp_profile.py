"""Per-kernel breakdown of BASELINE.json configs[2] (EIMP, batch 128, N = 2000 -> pruned, 9 iterations) through the library's
launch spans.  usage: python tools/eimp_profile.py"""
import json
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from imp_release_b200 import AdaGMN, ops  # noqa: E402
from oracle import synth  # noqa: E402

dev = torch.device('cuda')
B = 128
net = AdaGMN(bench.model_config(9))
net.load_state_dict(synth.make_state_dict('AdaGMN', 9, seed=7, bin_score=8.0))
net = net.to(dev).eval()
data = {k: v.to(dev) for k, v in synth.make_pair_batch(seed=2, batch=B, n0=2000, n1=2000).items()}
with torch.no_grad():
    ms = bench._timed(lambda: net(data), 2, 3)
    ops.PROFILE = {}
    net(data)
    torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
rows = sorted(((sum(a.elapsed_time(b) for a, b, _ in v), k, len(v)) for k, v in prof.items()), reverse=True)
tot = sum(r[0] for r in rows)
res = {'ms_per_batch': ms, 'pairs_per_s': B / ms * 1e3, 'kernel_ms_total': tot,
       'kernels': {k: {'ms': round(t, 3), 'calls': n, 'share': round(t / tot, 3)} for t, k, n in rows}}
res['sinkhorn_calls_ms'] = [round(a.elapsed_time(b), 3) for a, b, _ in prof.get('sinkhorn', [])]
res['score_gemm_calls_ms'] = [round(a.elapsed_time(b), 3) for a, b, _ in prof.get('gemm_n2000_k256_b', [])] + [round(a.elapsed_time(b), 3) for k, v in prof.items() if k.startswith('gemm_n') and k.endswith('_b') and k != 'gemm_n2000_k256_b' for a, b, _ in v]
res['kept'] = [int(x) for x in net._kept[0].view(2, -1).max(dim=1).values.tolist()]
print(json.dumps(res, indent=1))
json.dump(res, open('gpurun_out/eimp_profile.json', 'w'), indent=1)
