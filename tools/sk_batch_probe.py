"""Sinkhorn scoring time vs batch size / dustbin score (is batch 128 twice batch 64?).  usage: python tools/sk_batch_probe.py"""
import sys, torch
sys.path.insert(0, '.')
from imp_release_b200 import ops, _lib
lib = _lib.load()
N = 2000
for B in (32, 64, 96, 128):
    for bin_score in (1.0, 8.0):
        dist = torch.randn(B, N, N, device='cuda') * (3 if bin_score == 1.0 else 0.3)
        bs = torch.tensor(bin_score, device='cuda')
        ws = ops.SinkhornWorkspace(B, N, N, 'cuda')
        for _ in range(2):
            ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
        e1.record(); torch.cuda.synchronize()
        lib.imp_set_profiling(1)
        ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)
        it = float(lib.imp_sinkhorn_iter_ms())
        lib.imp_set_profiling(0)
        print(f'B={B} bin={bin_score}: scoring {e0.elapsed_time(e1) / 3:.3f} ms ({e0.elapsed_time(e1) / 3 / B * 64:.3f} per 64), iteration kernel {it:.4f} ms', flush=True)
        del dist, ws
        torch.cuda.empty_cache()
