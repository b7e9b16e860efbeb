"""Host -> device staging of pair batches (at >= 1 k pairs/s per GPU input staging is the next
limiter after the kernels).  The reference moves every batch with a blocking ``.cuda()`` on the compute stream
(eval/eval_imp.py:63-70), so the 265 MB of descriptors of a 64-pair batch sit in front of the matcher for ~5 ms.
``PairFeeder`` double-buffers the device copies on a dedicated copy stream: while the matcher works on batch i the pinned
host tensors of batch i+1 are already in flight.  Pure plumbing (streams, events, ``copy_``): no arithmetic here.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


class PairFeeder:
    """``stage(host_batch)`` enqueues the H2D copies of one batch on the copy stream; ``next()`` hands the oldest staged
    batch to the current stream (which waits for its copies, not for the host).  The batch handed out stays reserved until
    the following ``next()`` / ``release()`` records, on the consumer's stream, the event its slot's next copy waits for; so
    with ``depth`` buffers at most ``depth - 1`` batches can be staged ahead while one is in use.  Steady state with the
    default depth 2: ``stage(b0); loop: d = next(); stage(b_next); model(d)``."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self._ready = [torch.cuda.Event() for _ in range(depth)]       # copies of the slot done
        self._free = [None] * depth                                    # consumers of the slot done (None = never used)
        self._passthrough: List[Dict[str, object]] = [dict() for _ in range(depth)]
        self._head = self._tail = self._count = 0
        self._in_use: Optional[int] = None

    def stage(self, host_batch: Dict[str, object]) -> None:
        """Tensors are copied (they should be pinned for the copy to be asynchronous); non-tensor entries and tensors
        that only carry a shape ('image0' / 'image1', nets/gms.py:162-164) are passed through untouched."""
        if self._count + (1 if self._in_use is not None else 0) >= self.depth:
            raise RuntimeError('PairFeeder: no free slot (staged batches + the batch in use fill all %d buffers); '
                               'call next() / release() first' % self.depth)
        i = self._tail
        slot = self._slots[i]
        with torch.cuda.stream(self.copy_stream):
            if self._free[i] is not None:
                self.copy_stream.wait_event(self._free[i])
            if slot is None:
                slot = self._slots[i] = {}
            self._passthrough[i] = {}
            for k, v in host_batch.items():
                if not torch.is_tensor(v) or k.startswith('image'):
                    self._passthrough[i][k] = v
                    continue
                buf = slot.get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                    buf = slot[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                buf.copy_(v, non_blocking=True)
            self._ready[i].record(self.copy_stream)
        self._tail = (i + 1) % self.depth
        self._count += 1

    def next(self) -> Dict[str, object]:
        if self._count == 0:
            raise RuntimeError('PairFeeder: nothing staged')
        cur = torch.cuda.current_stream(self.device)
        self.release()
        i = self._head
        cur.wait_event(self._ready[i])
        self._head = (i + 1) % self.depth
        self._count -= 1
        self._in_use = i
        out: Dict[str, object] = dict(self._slots[i])
        out.update(self._passthrough[i])
        return out

    def release(self) -> None:
        """Mark the batch handed out by the last ``next()`` as consumed by everything enqueued so far on the current
        stream (called implicitly by the following ``next()``)."""
        if self._in_use is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._free[self._in_use] = ev
            self._in_use = None

    @staticmethod
    def bytes_of(host_batch: Dict[str, object]) -> int:
        return sum(v.numel() * v.element_size() for k, v in host_batch.items() if torch.is_tensor(v) and not k.startswith('image'))
