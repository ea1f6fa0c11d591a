"""Host->device staging for the training loop.

The reference moves every minibatch synchronously (`xs = (x.cuda() for x in xs)`,
src/drvish/train/__init__.py:42-43), so the copy of the 4096 x G fp32 count tile (82 MB at
config 3) serialises with the step.  ``PinnedBatchPrefetcher`` wraps any iterable of
``(x, labels, y)``-style batches, stages the large count tensor through pinned memory and issues
the copy of batch i+1 on a side stream while step i runs (two device buffers)."""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence

import torch


class PinnedBatchPrefetcher:
    def __init__(self, batches: Iterable[Sequence], device: torch.device):
        self.batches = batches
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._bufs = [None, None]
        self._events = [torch.cuda.Event(), torch.cuda.Event()]

    def _issue(self, slot: int, batch: Sequence):
        x = batch[0]
        if self._bufs[slot] is None or self._bufs[slot].shape != x.shape or self._bufs[slot].dtype != x.dtype:
            self._bufs[slot] = torch.empty(x.shape, dtype=x.dtype, device=self.device)
        if not x.is_cuda and not x.is_pinned():
            x = x.pin_memory()
        with torch.cuda.stream(self.stream):
            self._bufs[slot].copy_(x, non_blocking=True)
            rest = [t.to(self.device, non_blocking=True) if torch.is_tensor(t) and not t.is_cuda else t for t in batch[1:]]
            self._events[slot].record(self.stream)
        return [self._bufs[slot]] + rest

    def __iter__(self) -> Iterator[list]:
        it = iter(self.batches)
        try:
            nxt = self._issue(0, next(it))
        except StopIteration:
            return
        slot = 0
        while nxt is not None:
            cur, cur_slot = nxt, slot
            try:
                # the buffer of `slot ^ 1` was consumed by the step before last, which has completed
                # (SVI.step returns the Python float loss, i.e. it synchronises)
                slot ^= 1
                nxt = self._issue(slot, next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(self._events[cur_slot])
            yield cur
