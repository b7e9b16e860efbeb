"""Multi-GPU evaluation harness: image pairs are independent, so rank r of W processes (one per
B200) takes pairs r, r+W, ... with a full model replica; the only exchange is one gather of fixed-stride match
results (indices + scores) to rank 0.  No collective inside an iteration.  Works with any torch.distributed backend
(NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_indices(n_pairs: int, rank: int, world: int) -> List[int]:
    """Rank-strided partition: pair i goes to rank i % world."""
    return list(range(rank, n_pairs, world))


def gather_matches(indices0: torch.Tensor, mscores0: torch.Tensor, n_pairs: int, rank: int, world: int,
                   n_max: int) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """indices0 [local, n_max] int64 (-1 padded), mscores0 [local, n_max] fp32 for this rank's pairs in shard order.
    Returns ([n_pairs, n_max], [n_pairs, n_max]) in GLOBAL pair order on rank 0, None elsewhere."""
    local_max = (n_pairs + world - 1) // world
    dev = indices0.device
    pad_i = torch.full((local_max, n_max), -1, dtype=torch.int64, device=dev)
    pad_s = torch.zeros(local_max, n_max, dtype=torch.float32, device=dev)
    pad_i[:indices0.shape[0]] = indices0
    pad_s[:mscores0.shape[0]] = mscores0
    if world == 1 or not dist.is_initialized():
        return pad_i[:n_pairs], pad_s[:n_pairs]
    out_i = [torch.empty_like(pad_i) for _ in range(world)] if rank == 0 else None
    out_s = [torch.empty_like(pad_s) for _ in range(world)] if rank == 0 else None
    dist.gather(pad_i, out_i, dst=0)
    dist.gather(pad_s, out_s, dst=0)
    if rank != 0:
        return None
    full_i = torch.full((n_pairs, n_max), -1, dtype=torch.int64, device=dev)
    full_s = torch.zeros(n_pairs, n_max, dtype=torch.float32, device=dev)
    for r in range(world):
        ids = shard_indices(n_pairs, r, world)
        if ids:
            full_i[ids] = out_i[r][:len(ids)]
            full_s[ids] = out_s[r][:len(ids)]
    return full_i, full_s


def match_sharded(match_fn: Callable[[Sequence[int]], Tuple[torch.Tensor, torch.Tensor]], n_pairs: int, n_max: int,
                  rank: int, world: int):
    """Run ``match_fn(pair_ids) -> (indices0 [len, n_max], mscores0 [len, n_max])`` on this rank's shard and gather."""
    ids = shard_indices(n_pairs, rank, world)
    if ids:
        i0, s0 = match_fn(ids)
    else:
        dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
        i0 = torch.empty(0, n_max, dtype=torch.int64, device=dev)
        s0 = torch.empty(0, n_max, dtype=torch.float32, device=dev)
    return gather_matches(i0, s0, n_pairs, rank, world, n_max)


def evaluate_sharded(model, get_pair: Callable[[int], dict], n_pairs: int, n_max: int, rank: int, world: int,
                     slots: int = 4, p: float = 0.2, matcher=None):
    """BASELINE.json configs[3]: the one-pair-per-call evaluation of eval/eval_imp.py:155-173
    (``produce_matches(only_last=True)`` per pair, ragged keypoint counts) sharded over ranks.  Every rank replays bucketed
    CUDA graphs with ``slots`` pairs in flight (graphed.LatencyMatcher); ``get_pair(i)`` returns pair i's data dict
    (device tensors, or pinned host tensors -- they are staged on the slot's stream).  Rank 0 gets
    ([n_pairs, n_max] indices0, [n_pairs, n_max] mscores0), -1 / 0 padded.  Pass a ``matcher`` (LatencyMatcher) to keep the
    captured graphs across calls; otherwise one is built (and its graphs captured, ~50 ms per bucket and slot) per call."""
    from .graphed import LatencyMatcher
    dev = next(model.parameters()).device
    lm = matcher if matcher is not None else LatencyMatcher(model, slots=slots, p=p, only_last=True)

    timing = {}

    def match_fn(ids: Sequence[int]):
        import time
        i0 = torch.full((len(ids), n_max), -1, dtype=torch.int64, device=dev)
        s0 = torch.zeros(len(ids), n_max, dtype=torch.float32, device=dev)
        t0 = time.perf_counter()
        tickets = [lm.submit(get_pair(i)) for i in ids]
        timing['submit_s'] = time.perf_counter() - t0        # host time to enqueue everything (no synchronisation inside)
        for k, t in enumerate(tickets):
            out = lm.result(t)
            n = out['indices0'][-1].shape[1]
            i0[k, :n] = out['indices0'][-1][0]
            s0[k, :n] = out['mscores0'][-1][0]
        return i0, s0

    with torch.no_grad():
        out = match_sharded(match_fn, n_pairs, n_max, rank, world)
    evaluate_sharded.last_timing = timing
    return out
