"""CUDA-graph replay of one whole training step (forward + loss + backward + optimizer + metrics).

One U-Net step is ~160 kernel launches; captured once into a CUDA graph they replay without any host work or
launch gaps.  Everything launched through the C ABI is capturable: kernels are enqueued on torch's current
stream (the capture stream), tensor maps travel by value inside kernel parameters, and every scratch tensor
comes from the graph's private memory pool, so addresses are stable across replays.

    step = GraphedTrainStep(model, loss_fn, optimizer, X_example, y_example, metrics=agg)
    for X, y in loader:                       # X [N,3,H,W] fp32, y [N,H,W] or [N,1,H,W] integer labels
        loss = step(X, y)                     # device scalar; call loss.item() only when a value is needed

Constraints: fixed batch shape; the optimizer must be capturable (e.g. ``torch.optim.AdamW(..., fused=True)``
or ``capturable=True``); single process (data parallelism uses the eager path).  This is an addition for
throughput; the reference-compatible ``train_loop`` keeps working without it.
"""
from __future__ import annotations

import torch


class GraphedStep:
    """Generic form: ``step_fn(*static_inputs) -> device scalar`` captured once; ``__call__(*inputs)`` copies the new inputs
    into the static tensors and replays.  Used for the other model families (reconstruction: loss(model(x), x); prompt
    model: model(x, heatmap)), where the step is not ``loss_fn(model(x), y)``.

        step = GraphedStep(lambda x: train_step(x), [x_example], models=[model], warmup=2)
    """

    def __init__(self, step_fn, example_inputs, models=(), optimizer=None, warmup: int = 2):
        dev = example_inputs[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedStep needs CUDA inputs")
        self.step_fn, self.models = step_fn, list(models)
        self.static = [t.contiguous().clone() for t in example_inputs]
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                step_fn(*self.static)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = step_fn(*self.static)
        torch.cuda.synchronize(dev)
        self._engines = [getattr(m, "_engine", None) for m in self.models]

    def __call__(self, *inputs):
        if any(getattr(m, "_engine", None) is not e for m, e in zip(self.models, self._engines)):
            raise RuntimeError("this GraphedStep was captured for buffers a model no longer owns; capture a new one")
        for dst, src in zip(self.static, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for m in self.models:           # parameters changed inside the graph without bumping their version counters
            eng = getattr(m, "_engine", None)
            if eng is not None:
                for pool in eng.plans.values():
                    for plan in pool:
                        plan._pack_versions = None
        return self.out


class GraphedTrainStep:
    def __init__(self, model, loss_fn, optimizer, example_x: torch.Tensor, example_y: torch.Tensor, metrics=None,
                 warmup: int = 3):
        params = list(model.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device")
        self.model, self.loss_fn, self.optimizer, self.metrics = model, loss_fn, optimizer, metrics
        self.x = example_x.to(dev, dtype=torch.float32).contiguous().clone()
        y = example_y.to(dev)
        if y.dim() == 4:
            y = y[:, 0]
        self.y = y.long().contiguous().clone()
        self.loss = None
        self._engine = None
        model.train()
        # warm-up on a side stream (allocator pools, cudaFuncSetAttribute, optimizer state) as torch recommends
        optimizer.zero_grad(set_to_none=True)      # stale .grad tensors would be accumulated into by the warm-up step
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step()
        torch.cuda.synchronize(dev)
        # the graph replays into the buffers of THIS engine (activations, operand packs, scratch): keep it alive and refuse
        # to replay once the model has dropped it (model.to(...) / .float() re-create the engine)
        self._engine = getattr(model, "_engine", None)
        self._param_ptrs = tuple(p.data_ptr() for p in params)

    def _eager_step(self):
        pred = self.model(self.x)
        loss = self.loss_fn(pred, self.y)
        loss.backward()
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        if self.metrics is not None:
            self.metrics.accumulate(pred.detach(), self.y)
        return loss.detach()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if getattr(self.model, "_engine", None) is not self._engine or \
                tuple(p.data_ptr() for p in self.model.parameters()) != self._param_ptrs:
            raise RuntimeError("this GraphedTrainStep was captured for buffers the model no longer owns (the model was "
                               "moved / re-typed after the capture); capture a new one")
        self.x.copy_(x, non_blocking=True)
        if y.dim() == 4:
            y = y[:, 0]
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        # parameters were updated inside the graph without bumping their version counters: force the next EAGER
        # forward (e.g. evaluation) to re-pack the operand copies of the weights
        eng = getattr(self.model, "_engine", None)
        if eng is not None:
            for pool in eng.plans.values():
                for plan in pool:
                    plan._pack_versions = None
        return self.loss


class GraphedAccumulation:
    """Gradient accumulation (utils/training.py:49-56; the reference trains with micro-batch 2 x 32 accumulation steps,
    unet/unet.ipynb:41-42,64) as two CUDA graphs:

      micro graph   forward + loss + backward of ONE micro-batch; its gradients (views of the engine's flat buffer, taken
                    with ``torch.autograd.grad`` so that no per-parameter AccumulateGrad kernels run) are added to a
                    persistent flat accumulator with one multi-tensor add;
      step graph    optimizer.step() on ``p.grad`` (= views of the accumulator) and zeroing of the accumulator.

    ``micro(x, y)`` returns the micro-batch loss (device scalar); ``step()`` applies the update."""

    def __init__(self, model, loss_fn, optimizer, example_x, example_y, accumulation_steps: int):
        self.model, self.loss_fn, self.optimizer, self.k = model, loss_fn, optimizer, accumulation_steps
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        self.x = example_x.to(dev, dtype=torch.float32).contiguous().clone()
        y = example_y.to(dev)
        if y.dim() == 4:
            y = y[:, 0]
        self.y = y.long().contiguous().clone()
        total = sum(p.numel() for p in self.params)
        self.accum = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.accum[off:off + p.numel()].view(p.shape))
            off += p.numel()
        self.link()
        model.train()
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            self._micro()                       # warm-up (allocator pools, lazily built backward plan)
            self.accum.zero_()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        # The micro graph must NOT contain the weight (re)pack -- parameters change only in the step graph -- and the step
        # graph must END with it: the warm-up above has just packed (so the capture below finds nothing to do), and the
        # plan the capture runs on is identified by its generation counter.
        eng = model._engine
        plans = [pl for pool in eng.plans.values() for pl in pool]
        before = [pl.generation for pl in plans]
        self.micro_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.micro_graph):
            self.loss = self._micro()
        used = [pl for pl, g0 in zip(plans, before) if pl.generation != g0]
        if len(used) != 1:
            raise RuntimeError("could not identify the launch plan of the captured micro-batch step")
        self.plan = used[0]
        self.step_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.step_graph, pool=self.micro_graph.pool()):
            optimizer.step()
            self.accum.zero_()
            self.plan._pack_versions = None
            self.plan.pack_weights()                      # operand packs of the updated parameters, once per optimizer step
        self.accum.zero_()
        torch.cuda.synchronize(dev)
        self._engine = getattr(model, "_engine", None)
        self._param_ptrs = tuple(p.data_ptr() for p in self.params)

    def link(self):
        """``p.grad`` = views of the accumulator (what the optimizer reads; eager fall-back steps accumulate into them too)."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def unlink(self):
        for p in self.params:
            p.grad = None

    def sync_packs(self):
        """After an EAGER optimizer step (the first accumulation group of an epoch, a ragged last batch): the micro graph
        carries no weight pack, so the captured plan's operand packs are refreshed here."""
        with torch.cuda.device(self.params[0].device):
            self.plan._pack_versions = None
            self.plan.pack_weights()

    def _micro(self):
        pred = self.model(self.x)
        loss = self.loss_fn(pred, self.y)
        grads = torch.autograd.grad(loss / self.k, self.params)
        torch._foreach_add_(self.views, list(grads))
        return loss.detach()

    def _check(self):
        if getattr(self.model, "_engine", None) is not self._engine or \
                tuple(p.data_ptr() for p in self.params) != self._param_ptrs:
            raise RuntimeError("this GraphedAccumulation was captured for buffers the model no longer owns; capture a new one")

    def micro(self, x, y):
        self._check()
        self.x.copy_(x, non_blocking=True)
        if y.dim() == 4:
            y = y[:, 0]
        self.y.copy_(y, non_blocking=True)
        self.micro_graph.replay()
        return self.loss

    def step(self):
        self.step_graph.replay()
        eng = getattr(self.model, "_engine", None)
        if eng is not None:
            for pool in eng.plans.values():
                for plan in pool:
                    if plan is not self.plan:             # the captured plan was re-packed by the step graph itself
                        plan._pack_versions = None
