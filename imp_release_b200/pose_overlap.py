"""Overlap of the host-side pose / RANSAC step with the GPU matcher.

In the reference's evaluation loop every pair is matched on the GPU, its matches are copied to the host with blocking
``.cpu().numpy()`` calls and ``estimate_pose`` (cv2.findEssentialMat + recoverPose, eval/pose_estimation.py:92-115,
called from eval/eval_imp.py:167-173 and eval/matching.py:84) runs while the GPU idles -- once matching a pair takes a few
milliseconds the RANSAC (10-100 ms) is the whole wall clock.  ``PoseOverlap`` keeps the pose code untouched and on the host
but takes it off the GPU's critical path: the match indices of pair i are copied device -> pinned host memory
asynchronously on the caller's stream, a CUDA event marks the copy, and a worker thread waits for the event and runs the
caller's pose function while the main thread already enqueues pair i+1.  Results come back in submission order.

Pure plumbing: no arithmetic of the hot path lives here, and the pose function is whatever the caller passes (the
reference's ``estimate_pose`` in eval_imp.py).  Works with CPU tensors too (no event, the copy is immediate), which is how
the CPU tests drive it.
"""
from __future__ import annotations

from collections import deque
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any, Callable, Deque, List, Optional, Tuple

import torch


class PoseOverlap:
    """``submit(indices0, mscores0, fn, *args)`` -> Future of ``fn(indices0_np, mscores0_np, *args)``.

    ``indices0`` / ``mscores0`` are the matcher's outputs for one pair (``[N0]`` int64 / fp32, typically CUDA tensors, e.g.
    ``out['indices0'][-1][0]``); they are staged into pinned host buffers without synchronising the stream.  At most
    ``max_pending`` submissions may be in flight (their pinned buffers are recycled); ``submit`` blocks on the oldest one
    beyond that, which bounds host memory and keeps the GPU at most ``max_pending`` pairs ahead of the RANSAC."""

    def __init__(self, workers: int = 4, max_pending: int = 16):
        self._pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix='imp-pose')
        self._pending: Deque[Future] = deque()
        self._max_pending = max_pending
        self._free: List[Tuple[torch.Tensor, torch.Tensor]] = []

    def _buffers(self, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        for k, (bi, bs) in enumerate(self._free):
            if bi.numel() >= n:
                return self._free.pop(k)
        pin = torch.cuda.is_available()
        return (torch.empty(max(n, 1), dtype=torch.int64, pin_memory=pin),
                torch.empty(max(n, 1), dtype=torch.float32, pin_memory=pin))

    def submit(self, indices0: torch.Tensor, mscores0: Optional[torch.Tensor], fn: Callable[..., Any], *args, **kwargs) -> Future:
        while len(self._pending) >= self._max_pending:
            self._pending.popleft().exception()        # wait (result or error stays in the Future for the caller)
        n = indices0.numel()
        bi, bs = self._buffers(n)
        bi[:n].copy_(indices0.reshape(-1), non_blocking=True)
        if mscores0 is not None:
            bs[:n].copy_(mscores0.reshape(-1), non_blocking=True)
        event = None
        if indices0.is_cuda:
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream(indices0.device))
        has_scores = mscores0 is not None

        def work():
            try:
                if event is not None:
                    event.synchronize()                # waits for the copies of THIS pair only, not for the stream
                idx = bi[:n].numpy().copy()
                sc = bs[:n].numpy().copy() if has_scores else None
            finally:
                self._free.append((bi, bs))
            return fn(idx, sc, *args, **kwargs)

        fut = self._pool.submit(work)
        self._pending.append(fut)
        return fut

    def drain(self) -> None:
        """Wait for everything submitted so far."""
        while self._pending:
            self._pending.popleft().exception()

    def shutdown(self) -> None:
        self.drain()
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.shutdown()
        return False
