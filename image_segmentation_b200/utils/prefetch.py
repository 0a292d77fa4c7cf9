"""Host-to-device input pipeline for the train / eval loops.

The reference's loop (utils/training.py:41-47) copies every batch with a blocking ``X.to(device)`` right before the
forward pass, so the copy engine and the SMs take turns.  ``DevicePrefetcher`` wraps any iterable of ``(X, y)`` host
batches and keeps ONE batch in flight: while step *i* computes, the copy of batch *i+1* (from pinned memory, on a side
stream) is already running; the consumer's stream waits on an event, never on the host.
"""
from typing import Iterable, Iterator, Optional, Tuple

import torch


class DevicePrefetcher:
    def __init__(self, batches: Iterable, device, label_dtype: Optional[torch.dtype] = torch.int64, pin: bool = True):
        self.batches = batches
        self.device = torch.device(device)
        self.label_dtype = label_dtype
        self.pin = pin
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.batches)

    def _stage(self, batch) -> Tuple[torch.Tensor, torch.Tensor, torch.cuda.Event]:
        X, y = batch
        if self.pin:
            X = X if X.is_pinned() or X.is_cuda else X.pin_memory()
            y = y if y.is_pinned() or y.is_cuda else y.pin_memory()
        with torch.cuda.stream(self.stream):
            Xd = X.to(self.device, non_blocking=True)
            yd = y.to(self.device, non_blocking=True)
            if self.label_dtype is not None and yd.dtype != self.label_dtype:
                yd = yd.to(self.label_dtype)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return Xd, yd, ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        it = iter(self.batches)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            Xd, yd, ev = nxt
            try:
                nxt = self._stage(next(it))            # batch i+1 starts copying before batch i is consumed
            except StopIteration:
                nxt = None
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            Xd.record_stream(cur)                      # allocated on the side stream, consumed on the compute stream
            yd.record_stream(cur)
            yield Xd, yd


class AsyncScalarReader:
    """Device-to-host read-back of per-step scalars (the logged loss) without stalling the launch queue.

    The reference calls ``loss.item()`` right after ``optimizer.step()`` (utils/training.py:58), which blocks the host
    until the step has finished; the GPU then idles while the host enqueues the next step.  Here every scalar is copied
    to pinned memory asynchronously and read ONE step later: each step's value still reaches the host, in order, and
    sums / averages are unchanged."""

    def __init__(self, device, depth: int = 1):
        self.device = torch.device(device)
        self.depth = depth
        self.pending = []      # (pinned slot, event)
        self.free = []

    def push(self, value: torch.Tensor):
        slot = self.free.pop() if self.free else torch.empty(1, dtype=torch.float32).pin_memory()
        slot.copy_(value.detach().reshape(1).to(torch.float32), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.pending.append((slot, ev))

    def ready(self):
        """Values that can be read without waiting for the most recent ``depth`` steps."""
        out = []
        while len(self.pending) > self.depth:
            out.append(self._pop())
        return out

    def drain(self):
        out = []
        while self.pending:
            out.append(self._pop())
        return out

    def _pop(self) -> float:
        slot, ev = self.pending.pop(0)
        ev.synchronize()
        v = float(slot.item())
        self.free.append(slot)
        return v
