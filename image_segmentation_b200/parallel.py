"""Data-parallel training of the U-Net over NCCL / NVLink: one process per GPU, bucketed gradient
all-reduce overlapped with the remaining backward kernels.

The reference is single-process; the batch shards naturally, so the only exchange
step of the path is ONE sum-all-reduce of the 31 M fp32 gradients per optimiser step.  The engine
writes every parameter gradient of a backward pass into one flat buffer laid out in backward
completion order (head, up4 .. up1, down5 .. down1); as soon as a bucket of that buffer is final its
all-reduce is enqueued (``async_op=True``: NCCL's stream waits for the kernels enqueued so far and
then runs next to the dgrad/wgrad kernels that follow).  Averaging is folded into the upstream
gradient (``d logits / world``; every backward kernel is linear in it), so no extra pass over the
gradients exists.  BatchNorm batch statistics and Dice sums stay per rank (the reference has no SyncBN).

Difference from stock DistributedDataParallel: DDP's default ``broadcast_buffers=True`` re-broadcasts rank 0's
BatchNorm running statistics at every forward pass; here the buffers are broadcast once at construction and then
evolve per rank.  Call ``sync_buffers()`` before evaluating or checkpointing (``eval_loop`` followed by
``MetricsHistory.all_reduce`` must see the same running statistics on every rank).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class DataParallelUNet:
    """Attach bucketed gradient all-reduce to a ``unet`` instance (not a wrapper: the model object,
    its ``state_dict`` and the train loop stay exactly what they were)."""

    def __init__(self, model, process_group=None, bucket_mb: float = 25.0, broadcast_from: int = 0, compress=None,
                 exchange: str = "auto"):
        """``exchange``: "nccl" = bucketed ``ncclAllReduce`` (any backend torch.distributed offers, also gloo in the CPU
        tests); "nvls" = this package's two-shot multicast kernel (``unetk_nvls_allreduce_f32``: the gradients sit in
        symmetric memory and the NVSwitch reduces them, one launch after the last gradient); "auto" = nvls from four ranks
        on when the process group's devices support multicast, else nccl.
        ``bucket_mb``: minimum size of an all-reduce (a huge value = ONE all-reduce after the last gradient).
        ``compress="bf16"`` (opt-in, changes numerics): gradients travel as bf16 (half the bytes; the sum over ranks is
        formed in bf16 by NCCL) and are widened back to fp32 afterwards -- stock DDP's bf16 compression hook."""
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("torch.distributed must be initialised (backend nccl) before DataParallelUNet")
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.bucket_elems = max(1, int(bucket_mb * (1 << 20) / 4))
        if compress not in (None, "bf16"):
            raise ValueError("compress must be None or 'bf16'")
        self.compress = compress
        if exchange not in ("auto", "nccl", "nvls"):
            raise ValueError("exchange must be 'auto', 'nccl' or 'nvls'")
        self.exchange = exchange
        self._nvls = None            # (symmetric buffer, handle) once set up; False when unavailable
        import os
        self._trace = [] if os.environ.get("UNETK_DP_TRACE", "0") == "1" else None
        self._works: List = []
        self._sent = 0
        self._enabled = True
        model._grad_scale = 1.0 / self.world
        model._bucket_hook = self._on_bucket
        model._backward_done_hook = self._on_done
        if broadcast_from is not None:
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=broadcast_from, group=process_group)

    # engine callbacks ----------------------------------------------------------------------------
    # NVLS path ---------------------------------------------------------------------------------
    def _setup_nvls(self, plan):
        """Symmetric gradient buffer + multicast binding; collective (every rank gets here in its first backward pass).
        Any failure (no multicast support, no symmetric-memory backend) falls back to NCCL for 'auto' and raises for 'nvls'."""
        try:
            if self.compress is not None or not plan.flat_grad.is_cuda:
                raise RuntimeError("NVLS exchange needs CUDA tensors and fp32 gradients on the wire")
            import torch.distributed._symmetric_memory as symm_mem
            group = self.group if self.group is not None else dist.group.WORLD
            n = ((plan.grad_total + 3) // 4) * 4
            buf = symm_mem.empty(n, dtype=torch.float32, device=plan.flat_grad.device)
            hdl = symm_mem.rendezvous(buf, group)
            ok = bool(getattr(hdl, "has_multicast_support", False)) and int(hdl.multicast_ptr) != 0
            flag = torch.tensor([1 if ok else 0], device=buf.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)       # all ranks take the same path
            if int(flag.item()) != 1:
                raise RuntimeError("the devices of this process group do not support NVLink multicast")
            buf.zero_()
            self._nvls = (buf, hdl, n)
            # one exchange after the last gradient: the engine need not un-pack weight gradients block by block
            self.model._bucket_per_segment = False
        except Exception as e:
            if self.exchange == "nvls":
                raise RuntimeError(f"exchange='nvls' is not available here: {e}") from e
            self._nvls = False

    def _nvls_allreduce(self, plan):
        from . import _lib as L
        buf, hdl, n = self._nvls
        total = plan.grad_total
        trace = self._trace is not None and not torch.cuda.is_current_stream_capturing()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if trace else None

        def mark(i):
            if trace:
                ev[i].record()
        mark(0)
        buf[:total].copy_(plan.flat_grad)
        mark(1)
        hdl.barrier(channel=0)                       # every rank's gradients are in its symmetric buffer
        mark(2)
        L.nvls_allreduce(int(hdl.multicast_ptr), n, hdl.rank, hdl.world_size, 1.0)
        mark(3)
        hdl.barrier(channel=1)                       # every slice has been stored on every rank
        mark(4)
        plan.flat_grad.copy_(buf[:total])
        mark(5)
        if trace:
            self._trace.append(ev)

    def trace_summary(self):
        """Mean milliseconds of the pieces of the NVLS exchange over the traced (eager) steps: copy-in, barrier 0 (= how
        long this rank waits for the slowest one), multicast kernel, barrier 1, copy-out.  UNETK_DP_TRACE=1."""
        if not self._trace:
            return None
        torch.cuda.synchronize()
        names = ("copy_in", "wait_for_slowest_rank", "nvls_kernel", "barrier_after", "copy_out")
        rows = self._trace[len(self._trace) // 2:]          # skip warm-up
        return {nm: round(sum(e[i].elapsed_time(e[i + 1]) for e in rows) / len(rows), 4) for i, nm in enumerate(names)}

    def _use_nvls(self, plan) -> bool:
        if self.exchange == "nccl" or self.world == 1:
            return False
        if self.exchange == "auto" and self.world < 4:
            # two ranks exchange halves point to point; the switch reduction pays from four ranks on (measured on 2 x B200:
            # 21.31 ms/step with the multicast kernel, 21.07 with NCCL; a tie at 8 -- DESIGN.md section 6)
            self._nvls = False
            return False
        if self._nvls is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("run one eager data-parallel step before capturing a CUDA graph (the symmetric gradient "
                                   "buffer is set up in the first backward pass)")
            self._setup_nvls(plan)
        return bool(self._nvls)

    def _on_bucket(self, plan, end_offset: int):
        if not self._enabled or self.world == 1:
            return
        if self.exchange != "nccl" and self._nvls is not False and plan.flat_grad.is_cuda:
            if self._use_nvls(plan):
                return                               # one multicast all-reduce after the last gradient (_on_done)
        if getattr(self, "_plan_generation", None) != (id(plan), getattr(plan, "generation", 0)):
            # first bucket of a new backward pass: a previous pass that raised midway must not leave stale offsets / works
            self._plan_generation = (id(plan), getattr(plan, "generation", 0))
            self._works = []
            self._sent = 0
        if end_offset - self._sent >= self.bucket_elems:
            self._launch(plan, end_offset)

    def _launch(self, plan, end_offset):
        chunk = plan.flat_grad[self._sent:end_offset]
        if self.compress == "bf16":
            wire = chunk.to(torch.bfloat16)
            self._works.append((dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group, async_op=True), chunk, wire))
        else:
            self._works.append((dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True), None, None))
        self._sent = end_offset

    def _on_done(self, plan):
        if self._enabled and self.world > 1 and self.exchange != "nccl" and plan.flat_grad.is_cuda and self._use_nvls(plan):
            self._nvls_allreduce(plan)
            self._works, self._sent, self._plan_generation = [], 0, None
            return
        if self._enabled and self.world > 1:
            if self._sent < plan.grad_total:
                self._launch(plan, plan.grad_total)
            for w, chunk, wire in self._works:
                w.wait()          # current stream waits for NCCL; the host does not block
                if wire is not None:
                    chunk.copy_(wire)
        self._works = []
        self._sent = 0
        self._plan_generation = None

    def sync_buffers(self, src: int = 0):
        """Broadcast rank ``src``'s BatchNorm running statistics (what DDP's broadcast_buffers does every forward)."""
        for t in self.model.buffers():
            dist.broadcast(t.data, src=src, group=self.group)

    # Gradient accumulation (utils/training.py:49-56): every micro-batch is reduced.  The flat buffer holds only
    # the CURRENT backward's gradients (autograd adds them to .grad afterwards), so skipping the reduction for all
    # but the last micro-batch -- DDP's no_sync -- would lose the earlier ones; sum-of-averages is the same number
    # and each 124 MB all-reduce hides behind the backward kernels.

    def detach(self):
        for name in ("_grad_scale", "_bucket_hook", "_backward_done_hook", "_bucket_per_segment"):
            if hasattr(self.model, name):
                delattr(self.model, name)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous batch shard of rank ``rank`` (the path partitions over the batch dimension only)."""
    n = x.shape[0]
    if n % world:
        raise ValueError(f"batch {n} does not divide over {world} ranks")
    per = n // world
    return x[rank * per:(rank + 1) * per]
