"""Host-to-device input pipeline for the train / eval loops.

The reference's loop (utils/training.py:41-47) copies every batch with a blocking ``X.to(device)`` right before the
forward pass, so the copy engine and the SMs take turns.  ``DevicePrefetcher`` wraps any iterable of ``(X, y)`` host
batches and keeps ONE batch in flight: while step *i* computes, the copy of batch *i+1* (from pinned memory, on a side
stream) is already running; the consumer's stream waits on an event, never on the host.
"""
from typing import Iterable, Iterator, Optional, Tuple

import torch


class DevicePrefetcher:
    """Two device slots per tensor, filled alternately on a side stream; no allocation in steady state (a fresh device
    tensor per batch would go through cudaMalloc / deferred frees of the caching allocator and make step times jump)."""

    def __init__(self, batches: Iterable, device, label_dtype: Optional[torch.dtype] = torch.int64, pin: bool = True):
        self.batches = batches
        self.device = torch.device(device)
        self.label_dtype = label_dtype
        self.pin = pin
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [None, None]          # per slot: (X_dev, y_raw_dev, y_dev)

    def __len__(self):
        return len(self.batches)

    def _slot(self, k, X, y):
        cur = self._slots[k]
        ydt = self.label_dtype if self.label_dtype is not None else y.dtype
        if cur is None or cur[0].shape != X.shape or cur[0].dtype != X.dtype or cur[1].shape != y.shape \
                or cur[1].dtype != y.dtype or cur[2].dtype != ydt:
            Xd = torch.empty(X.shape, dtype=X.dtype, device=self.device)
            yr = torch.empty(y.shape, dtype=y.dtype, device=self.device)
            yd = yr if ydt == y.dtype else torch.empty(y.shape, dtype=ydt, device=self.device)
            cur = self._slots[k] = (Xd, yr, yd)
        return cur

    def _stage(self, batch, k, after: Optional[torch.cuda.Event]):
        X, y = batch
        if X.is_cuda and y.is_cuda:       # already resident: nothing to copy
            yd = y if (self.label_dtype is None or y.dtype == self.label_dtype) else y.to(self.label_dtype)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            return X, yd, ev
        if self.pin:
            X = X if X.is_pinned() else X.pin_memory()
            y = y if y.is_pinned() else y.pin_memory()
        Xd, yr, yd = self._slot(k, X, y)
        with torch.cuda.stream(self.stream):
            if after is not None:
                self.stream.wait_event(after)            # the step that last read this slot has been enqueued and is done
            Xd.copy_(X, non_blocking=True)
            yr.copy_(y, non_blocking=True)
            if yd is not yr:
                yd.copy_(yr)                             # dtype widening on the device (uint8 label maps -> int64)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return Xd, yd, ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        it = iter(self.batches)
        try:
            nxt = self._stage(next(it), 0, None)
        except StopIteration:
            return
        k = 0
        while nxt is not None:
            Xd, yd, ev = nxt
            cur = torch.cuda.current_stream(self.device)
            # everything enqueued so far (the step on batch k-1, which lives in the other slot) precedes this event;
            # the refill of that slot waits for it
            done_prev = torch.cuda.Event()
            done_prev.record(cur)
            try:
                nxt = self._stage(next(it), (k + 1) % 2, done_prev)   # batch k+1 starts copying before batch k is consumed
            except StopIteration:
                nxt = None
            cur.wait_event(ev)
            yield Xd, yd
            k += 1


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
