"""Drop-in for the reference's ``utils/training.py``: ``train_loop`` (:18-64), ``eval_loop`` (:67-121),
``trainReconstruction`` (:123-151), ``train_loop_prompt`` (:153-199), ``evalReconstruction`` (:202-239),
``eval_loop_prompt`` (:242-296), ``start_prompt`` (:299-450) and ``start`` (:453-617) with the same signatures, printed
lines and checkpoint dictionary keys.  Model, loss and metrics objects are passed in, exactly as in the
reference; with this package's ``unet`` / ``WeightedDiceCELoss`` / ``MetricsHistory`` every step runs on
the CUDA kernels, and the loops themselves only sequence work (host code stays Python).

Changes that do not alter results:
  * batches that already have the target resolution skip the per-image resize loop;
  * the loss value is read back once per optimiser step (as the reference does) -- no other syncs.
"""
import os
import sys

import torch

from .. import _lib as L
from .MetricsHistory import MetricsHistory
from .prefetch import AsyncScalarReader, DevicePrefetcher
from .weighted_loss import WeightedDiceCELoss
from .utils import process_batch_forward, process_batch_reverse

try:  # progress bars are optional plumbing
    from tqdm.auto import tqdm as _tqdm
except Exception:  # pragma: no cover
    _tqdm = None


def _progress(iterable, **kw):
    # the reference always shows tqdm bars; they are shown here as well unless stdout is not a terminal (logs, tests)
    # or UNETK_NO_TQDM=1
    quiet = os.environ.get("UNETK_NO_TQDM")
    if quiet is None:
        quiet = "0" if sys.stderr.isatty() else "1"
    if _tqdm is None or quiet == "1":
        class _P:
            def __init__(self, it):
                self._it = it

            def __iter__(self):
                return iter(self._it)

            def set_postfix(self, *a, **k):
                pass
        return _P(iterable)
    return _tqdm(iterable, **kw)


def _graph_eligible(model, loss_fn, optimizer, accumulation_steps, scheduler, device) -> bool:
    """Whole-step CUDA-graph replay inside train_loop needs: this package's unet + fused loss on a CUDA device and an
    optimizer whose step is capturable (``torch.optim.AdamW(..., capturable=True)``).  Gradient accumulation
    (``accumulation_steps > 1``, the reference's own configuration: micro-batch 2 x 32) replays a micro-batch graph and an
    optimizer graph (``GraphedAccumulation``).  A learning-rate scheduler is fine when every ``lr`` is a DEVICE tensor
    (``AdamW(lr=torch.tensor(1e-3, device=...))``: schedulers then update it in place and the captured step reads the new
    value); with python-float learning rates a scheduler keeps the loop eager.  ``UNETK_TRAIN_GRAPH=0`` turns it off."""
    from ..unet.unet import unet as _unet
    from .weighted_loss import WeightedDiceCELoss as _Loss
    if os.environ.get("UNETK_TRAIN_GRAPH", "1") != "1" or torch.device(device).type != "cuda":
        return False
    if not isinstance(model, _unet) or not isinstance(loss_fn, _Loss):
        return False
    if scheduler is not None and not all(torch.is_tensor(g["lr"]) and g["lr"].is_cuda for g in optimizer.param_groups):
        return False
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        return False
    return all(g.get("capturable", False) for g in optimizer.param_groups)


def _invalidate_weight_packs(model):
    """Host-side bookkeeping done during an aborted capture (weight-pack versions) must not survive it: the captured
    pack kernel never ran."""
    eng = getattr(model, "_engine", None)
    if eng is not None:
        for pool in eng.plans.values():
            for plan in pool:
                plan._pack_versions = None


def _graph_signature(model, loss_fn, optimizer, X, y):
    """Everything a captured step bakes in: batch shape, the addresses of parameters and optimizer state, the optimizer's
    hyper-parameters (python floats are frozen into the capture; tensors are read at replay time, so only their address
    counts), the loss configuration and the model's precision / kernel selection.  Any difference => re-capture."""
    def norm(v):
        if torch.is_tensor(v):
            return ("tensor", v.data_ptr())
        if isinstance(v, (list, tuple)):
            return tuple(norm(e) for e in v)
        return v
    params = [p for g in optimizer.param_groups for p in g["params"]]
    state = tuple(t.data_ptr() for p in params for _, t in sorted(optimizer.state.get(p, {}).items()) if torch.is_tensor(t))
    hyper = tuple(tuple((k, norm(v)) for k, v in sorted(g.items()) if k != "params") for g in optimizer.param_groups)
    cw = getattr(loss_fn, "class_weights", None)
    loss_cfg = (getattr(loss_fn, "dice_weight", None), getattr(loss_fn, "ce_weight", None),
                getattr(loss_fn, "ignore_index", None), getattr(loss_fn, "smooth_dice", None),
                None if cw is None else (cw.data_ptr(), cw._version))
    return (tuple(X.shape), tuple(y.shape), tuple(p.data_ptr() for p in model.parameters()),
            tuple(p.data_ptr() for p in params), state, hyper, loss_cfg, model.precision, model.conv_algo, model.training)


def _graph_for(model, loss_fn, optimizer, X, y, accumulation_steps=1):
    """Captured step for this (model, loss, optimizer, batch shape); cached on the model across epochs.

    The cache entry holds STRONG references to the loss, the optimizer and the engine whose buffers the graph replays
    into, and is compared by identity (``is``) plus ``_graph_signature``: a moved model (``unet._apply`` drops the engine
    and this cache), a restored optimizer state (``load_state_dict`` replaces the state tensors), a changed learning
    rate / weight decay, or a different loss object all lead to a fresh capture instead of a replay into stale memory."""
    from .graph import GraphedAccumulation, GraphedTrainStep
    sig = _graph_signature(model, loss_fn, optimizer, X, y) + (accumulation_steps,)
    cached = getattr(model, "_train_graph", None)
    if cached is not None:
        c_loss, c_opt, c_engine, c_sig, g = cached
        if c_loss is loss_fn and c_opt is optimizer and c_engine is model._engine and c_sig == sig:
            return g
        model._train_graph = None        # stale: release the old graph (and its private memory pool) before re-capturing
        del cached, g
    try:
        if accumulation_steps > 1:
            g = GraphedAccumulation(model, loss_fn, optimizer, X, y, accumulation_steps)
        else:
            g = GraphedTrainStep(model, loss_fn, optimizer, X, y, metrics=None, warmup=0)
    except Exception as e:  # pragma: no cover - capture is an optimisation, never a requirement
        print(f"[train_loop] CUDA graph capture unavailable, staying eager: {e!r}")
        _invalidate_weight_packs(model)
        optimizer.zero_grad(set_to_none=True)
        return None
    # the signature is taken AFTER the capture: its warm-up may have created optimizer state
    model._train_graph = (loss_fn, optimizer, model._engine,
                          _graph_signature(model, loss_fn, optimizer, X, y) + (accumulation_steps,), g)
    return g


def train_loop(dataloader, model, loss_fn, optimizer, accumulation_steps, device, scheduler=None, target_size=None):
    """One epoch of training with gradient accumulation; returns the mean logged loss per optimiser step."""
    model.train()
    total_loss, processed_batches = 0.0, 0
    num_batches = len(dataloader)
    optimizer.zero_grad()
    def host_batches():
        for X, y in dataloader:
            if target_size is not None:
                X, _ = process_batch_forward(X, target_size=target_size)
                y, _ = process_batch_forward(y, target_size=target_size, interpolation="nearest")
            yield X, y

    # the copy of batch i+1 overlaps the step on batch i (utils/prefetch.py); results are unchanged
    if torch.device(device).type == "cuda":
        batches = DevicePrefetcher(host_batches(), device)
    else:
        batches = ((X.to(device), y.to(device).long()) for X, y in host_batches())
    pbar = _progress(enumerate(batches), total=num_batches, desc="Training")
    # the logged loss of every optimiser step is read back one step late (AsyncScalarReader): same values, same order
    reader = AsyncScalarReader(device) if torch.device(device).type == "cuda" else None
    graph_ok = _graph_eligible(model, loss_fn, optimizer, accumulation_steps, scheduler, device)
    graphed = None
    accum = accumulation_steps > 1
    try:
        for batch_idx, (X, y) in pbar:
            step_now = (batch_idx + 1) % accumulation_steps == 0 or (batch_idx + 1) == num_batches
            # the first optimiser step of the epoch runs eagerly and doubles as warm-up; from then on the whole step
            # (forward + loss + backward [+ optimizer]) replays as CUDA graphs -- same kernels, no launch gaps
            if graph_ok and graphed is None and batch_idx >= accumulation_steps and batch_idx % accumulation_steps == 0:
                graphed = _graph_for(model, loss_fn, optimizer, X, y, accumulation_steps)
                graph_ok = graphed is not None
                if graphed is not None and accum:
                    graphed.link()
                    graphed.sync_packs()                  # the eager optimizer step just before changed the parameters
            use_graph = graphed is not None and tuple(X.shape) == tuple(graphed.x.shape)
            if use_graph and not accum:
                loss = graphed(X, y)                     # optimizer.step() and zero_grad() are part of the graph
                if scheduler:
                    scheduler.step()
            elif use_graph:
                loss = graphed.micro(X, y)
                if step_now:
                    graphed.step()
                    if scheduler:
                        scheduler.step()
            else:
                pred = model(X)
                loss = loss_fn(pred, y.squeeze(1))
                (loss / accumulation_steps).backward()
                # drop the autograd graph now: a graph kept alive by `loss` would pin its AccumulateGrad nodes to this
                # stream and invalidate the CUDA-graph capture of the next step
                loss = loss.detach()
                del pred
                if step_now:
                    optimizer.step()
                    if scheduler:
                        scheduler.step()
                    # with a live accumulation graph p.grad are views of its accumulator: zero them in place
                    optimizer.zero_grad(set_to_none=not (graphed is not None and accum))
                    if graphed is not None and accum:
                        graphed.sync_packs()
            if step_now:
                if reader is not None:
                    reader.push(loss)
                    values = reader.ready()
                else:
                    values = [loss.item()]
                for value in values:
                    total_loss += value
                    processed_batches += 1
                    pbar.set_postfix({'loss': value, 'lr': optimizer.param_groups[0]['lr']})
    finally:
        if graphed is not None and accum:
            graphed.unlink()                              # like the reference, gradients are None after the loop
    if reader is not None:
        for value in reader.drain():
            total_loss += value
            processed_batches += 1
    # out-of-range labels (the reference raises from scatter_, utils/weighted_loss.py:58): graph replays never run the
    # Python side of the loss again, so the device flag is read here, once per epoch, after the queue has drained
    status = getattr(loss_fn, "_status", None)
    if status is not None and torch.device(device).type == "cuda":
        status.publish()
        status.check(True, type(loss_fn).__name__)
    avg_loss = total_loss / processed_batches if processed_batches > 0 else 0
    print(f"Training Avg loss (per effective batch): {avg_loss:>8f}")
    return avg_loss


class _EvalTail:
    """Fused evaluation tail for one epoch (utils/training.py:93-101 + utils/utils.py:101-115): every batch is one
    ``unetk_eval_loss_metrics`` call that resizes the logits of all its images back to their original sizes on the fly
    and reduces the per-image Dice+CE losses and the confusion counts on the device.  Nothing is read back until
    ``finish()``."""

    def __init__(self, loss_fn, agg, device):
        self.loss_fn, self.agg, self.device = loss_fn, agg, torch.device(device)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.per_image = []
        cw = loss_fn.class_weights
        self.cw = None if cw is None else cw.detach().to(device=self.device, dtype=torch.float32).contiguous()

    @staticmethod
    def supports(loss_fn, agg, preds):
        return (isinstance(loss_fn, WeightedDiceCELoss) and isinstance(agg, MetricsHistory) and preds.is_cuda
                and preds.dim() == 4 and preds.shape[1] == agg.get_num_classes() and preds.shape[1] <= 8)

    def batch(self, preds, meta_list, labels):
        n, c, th, tw = preds.shape
        if len(labels) != n or len(meta_list) != n:
            raise ValueError(f"{n} predictions, {len(meta_list)} metas, {len(labels)} labels")
        flat = []
        for lab, m in zip(labels, meta_list):
            oh, ow = m["original_size"]
            if lab.numel() != oh * ow:
                raise ValueError(f"Shape mismatch: label {tuple(lab.shape)} vs original size {(oh, ow)}")
            flat.append(lab.reshape(-1))
        if not all(f.dtype == torch.uint8 for f in flat):
            flat = [f.long() for f in flat]
        packed = torch.cat(flat)
        if not packed.is_cuda:
            packed = packed.pin_memory().to(self.device, non_blocking=True)
        lg = preds.detach()
        lg = lg if (lg.dtype == torch.float32 and lg.is_contiguous()) else lg.float().contiguous()
        fn = self.loss_fn
        ign = fn.ignore_index
        if self.cw is not None and self.cw.numel() != c:
            raise RuntimeError(f"weight tensor should be defined for all {c} classes, got {self.cw.numel()}")
        with torch.cuda.device(self.device):
            table, _, mx = L.eval_image_table(meta_list, self.device)
            accum = torch.zeros(n, 3 * c + 2, dtype=torch.float64, device=self.device)
            per = torch.empty(n, dtype=torch.float32, device=self.device)
            counts, status = self.agg._device_counters(self.device)
            args = L.EvalArgs(lg.data_ptr(), n, c, th, tw, table.data_ptr(), mx, packed.data_ptr(),
                              L.UNETK_U8 if packed.dtype == torch.uint8 else L.UNETK_I64, L.ptr(self.cw),
                              0 if ign is None else 1, 0 if ign is None else int(ign), float(fn.dice_weight),
                              float(fn.ce_weight), float(fn.smooth_dice), accum.data_ptr(), per.data_ptr(),
                              self.loss_sum.data_ptr(), counts.data_ptr(), status.data_ptr())
            L.eval_loss_metrics(args)
        self.per_image.append(per)
        self._keep = (lg, packed, table, accum)          # alive until the next batch is enqueued on the same stream

    def finish(self):
        return float(self.loss_sum.item())               # the one sync of the epoch


def eval_loop(dataloader, model, loss_fn, device, target_size, agg):
    """Evaluation at each image's original resolution: mean loss, macro Dice and mIoU."""
    model.eval()
    num_images_processed = 0
    losses = []
    tail = None
    agg.reset()
    with torch.no_grad():
        for X, y in _progress(dataloader, desc="Eval"):
            X, meta_list = process_batch_forward(X, target_size=target_size)
            preds = model(X.to(device, non_blocking=True))
            if _EvalTail.supports(loss_fn, agg, preds):
                if tail is None:
                    tail = _EvalTail(loss_fn, agg, preds.device)
                tail.batch(preds, meta_list, list(y))
                num_images_processed += len(meta_list)
                continue
            # any other loss / metrics object: materialise the resized predictions (one launch) and call it per image
            preds = process_batch_reverse(preds, meta_list, interpolation='bilinear')
            for pred, label in zip(preds, y):
                label = label.to(device).long()
                losses.append(loss_fn(pred.unsqueeze(0), label.unsqueeze(0).squeeze(1)))
                agg.accumulate(pred, label)
                num_images_processed += 1
    total = tail.finish() if tail is not None else 0.0
    if losses:
        total += torch.stack([l.detach().double().reshape(()) for l in losses]).sum().item()
    avg_loss = total / num_images_processed if num_images_processed else float("nan")
    mean_dice, mean_iou, mean_acc = agg.compute_epoch_metrics()
    per_class_iou = agg.get_last_per_class_iou()
    print(f"\n--- Evaluation Complete ---")
    print(f"  Images Processed: {num_images_processed}")
    print(f"  Average Loss (Original Size): {avg_loss:>8f}")
    print(f"  Ignored Class : {agg.get_ignore_index()}")
    print(f"  Macro Avg Acc score: {mean_acc:>8f}")
    print(f"  Macro Avg Dice Score: {mean_dice:>8f}")
    print(f"  Mean IoU (mIoU): {mean_iou:>8f}")
    print(f"  --- Per-Class IoU ---")
    for c in range(agg.get_num_classes()):
        print(f"    Class {c}: {per_class_iou[c].item():>8f}")
    print("-" * 25)
    return avg_loss, mean_dice, mean_iou


def trainReconstruction(dataloader, model, loss_fn, optimizer, accumulation_steps, device=None):
    """One epoch of reconstruction pre-training (reference utils/training.py:123-151): ``loss_fn(model(X), X)`` with
    gradient accumulation; returns the mean of the per-batch losses.  The reference reads a module-global ``device``;
    here it defaults to the device of the model's parameters.  Per-batch losses are read back one step late
    (``AsyncScalarReader``): same values, same order, no pipeline stall."""
    import numpy as np
    device = torch.device(device) if device is not None else next(model.parameters()).device
    losses = []
    model.train()
    num_batches = len(dataloader)
    cuda = device.type == "cuda"
    reader = AsyncScalarReader(device, depth=2) if cuda else None
    for batch_idx, (X, _) in enumerate(_progress(dataloader, total=num_batches, desc="Training")):
        X = X.to(device, non_blocking=True)
        pred = model(X)
        loss = loss_fn(pred, X)
        if reader is not None:
            reader.push(loss)
            losses += reader.ready()
        else:
            losses.append(loss.item())
        (loss / accumulation_steps).backward()
        if (batch_idx + 1) % accumulation_steps == 0 or (batch_idx + 1) == num_batches:
            optimizer.step()
            optimizer.zero_grad()
    if reader is not None:
        losses += reader.drain()
    return np.mean(losses)


def evalReconstruction(dataloader, model, loss_fn, target_size, interpolation='bilinear', device=None):
    """Evaluation of a reconstruction model at each image's original size (reference utils/training.py:202-239);
    returns (total_loss / num_batches, mean per-image loss)."""
    import numpy as np
    device = torch.device(device) if device is not None else next(model.parameters()).device
    model.eval()
    num_batches = len(dataloader)
    per_image = []
    with torch.no_grad():
        for original_X, _ in _progress(dataloader, total=num_batches, desc="Evaluation"):
            resized_X, meta_list = process_batch_forward(original_X, target_size=target_size)
            pred = model(resized_X.to(device, non_blocking=True))
            pred = process_batch_reverse(pred, meta_list, interpolation=interpolation)
            for p, label in zip(pred, original_X):
                p = p.to(device).unsqueeze(0)
                label = label.to(device).unsqueeze(0)
                if label.shape[1] == 4 and label.ndim == 4:
                    label = label[:, :3, :, :]          # RGBA to RGB
                per_image.append(loss_fn(p, label.squeeze(1)).detach().double().reshape(()))
    vals = torch.stack(per_image).cpu().numpy() if per_image else np.zeros(0)     # one read-back for the whole epoch
    total_loss = float(vals.sum())
    return total_loss / num_batches, np.mean(vals)


def train_loop_prompt(dataloader, model, loss_fn, optimizer, accumulation_steps, device, scheduler=None, target_size=None):
    """One epoch of training of a prompt-based model, batches of (image, point-prompt heat-map, label)
    (reference utils/training.py:153-199)."""
    model.train()
    total_loss, processed_batches = 0.0, 0
    optimizer.zero_grad()
    num_batches = len(dataloader)
    cuda = torch.device(device).type == "cuda"
    reader = AsyncScalarReader(device) if cuda else None
    pbar = _progress(enumerate(dataloader), total=num_batches, desc="Training")
    for batch_idx, (X, p, y) in pbar:
        if target_size is not None:
            X, _ = process_batch_forward(X, target_size=target_size)
            p, _ = process_batch_forward(p, target_size=target_size)
            y, _ = process_batch_forward(y, target_size=target_size, interpolation="nearest")
        X, p, y = X.to(device, non_blocking=True), p.to(device, non_blocking=True), y.to(device, non_blocking=True).long()
        pred = model(X, p)
        loss = loss_fn(pred, y.squeeze(1))
        (loss / accumulation_steps).backward()
        loss = loss.detach()
        if (batch_idx + 1) % accumulation_steps == 0 or (batch_idx + 1) == num_batches:
            optimizer.step()
            if scheduler:
                scheduler.step()
            optimizer.zero_grad()
            if reader is not None:
                reader.push(loss)
                values = reader.ready()
            else:
                values = [loss.item()]
            for value in values:
                total_loss += value
                processed_batches += 1
                pbar.set_postfix({'loss': value, 'lr': optimizer.param_groups[0]['lr']})
    if reader is not None:
        for value in reader.drain():
            total_loss += value
            processed_batches += 1
    avg_loss = total_loss / processed_batches if processed_batches > 0 else 0
    print(f"Training Avg loss (per effective batch): {avg_loss:>8f}")
    return avg_loss


def eval_loop_prompt(dataloader, model, loss_fn, device, target_size, agg):
    """Evaluation of a prompt-based model at each image's original resolution (reference utils/training.py:242-296).
    Like the reference, ``agg`` is NOT reset here."""
    model.eval()
    num_images_processed = 0
    losses = []
    with torch.no_grad():
        for X, p, y in _progress(dataloader, desc="Eval"):
            X, meta_list = process_batch_forward(X, target_size=target_size)
            p, _ = process_batch_forward(p, target_size=target_size)
            preds = model(X.to(device, non_blocking=True), p.to(device, non_blocking=True))
            preds = process_batch_reverse(preds, meta_list, interpolation='bilinear')
            for pred, label in zip(preds, y):
                label = label.to(device).long()
                losses.append(loss_fn(pred.unsqueeze(0), label.unsqueeze(0).squeeze(1)).detach().double().reshape(()))
                agg.accumulate(pred, label)
                num_images_processed += 1
    total_loss = float(torch.stack(losses).sum().item()) if losses else 0.0
    avg_loss = total_loss / num_images_processed
    mean_dice, mean_iou, mean_acc = agg.compute_epoch_metrics()
    per_class_iou = agg.get_last_per_class_iou()
    print(f"\n--- Evaluation Complete ---")
    print(f"  Images Processed: {num_images_processed}")
    print(f"  Average Loss (Original Size): {avg_loss:>8f}")
    print(f"  Ignored Class : {agg.get_ignore_index()}")
    print(f"  Macro Avg Acc score: {mean_acc:>8f}")
    print(f"  Macro Avg Dice Score: {mean_dice:>8f}")
    print(f"  Mean IoU (mIoU): {mean_iou:>8f}")
    print(f"  --- Per-Class IoU ---")
    for c in range(4):
        print(f"    Class {c}: {per_class_iou[c].item():>8f}")
    print("-" * 25)
    return avg_loss, mean_dice, mean_iou


def start(model_save_dir, model_save_name, model, optimizer, train_dataloader, val_dataloader, accumulation_steps,
          device, train_loss_fn, val_loss_fn, target_size, scheduler=None, agg=None, load=True, save=True,
          num_classes=4, ignore_index=3, epochs=100, _prompt=False):
    """Train/evaluate for ``epochs`` epochs with best-mIoU checkpointing and resume (reference :453-617).

    Same printed lines, files and checkpoint keys as the reference.  Deliberate differences: the model is moved to
    ``device`` if it is not there yet (the reference expects the caller to have done it), and a missing ``agg`` is
    created instead of failing in ``eval_loop``."""
    dev = torch.device(device)
    p0 = next(model.parameters(), None)
    if p0 is not None and (p0.device.type != dev.type or (dev.index is not None and p0.device.index != dev.index)):
        model.to(device)
    if agg is None:
        agg = MetricsHistory(num_classes, ignore_index)
    best = {"best_dev_dice": -float("inf"), "best_dev_miou": -float("inf"), "best_dev_loss": float("inf")}
    start_epoch = 0
    path = os.path.join(model_save_dir, model_save_name)
    os.makedirs(model_save_dir, exist_ok=True)
    os.makedirs(os.path.join(model_save_dir, "metrics"), exist_ok=True)
    if load and os.path.isfile(path):
        print(f"Loading checkpoint from: {path}")
        # start_prompt() loads with pickle enabled: its checkpoints carry the MetricsHistory object (reference :351,:423)
        ckpt = torch.load(path, map_location=device, weights_only=not _prompt)
        model.load_state_dict(ckpt["model_state_dict"])
        print(" -> Model state loaded.")
        for key, obj, what in (("optimizer_state_dict", optimizer, "Optimizer"), ("scheduler_state_dict", scheduler, "Scheduler")):
            try:
                obj.load_state_dict(ckpt[key])
                print(f" -> {what} state loaded.")
            except Exception as e:  # same tolerance (and wording) as the reference: a missing scheduler lands here too
                print(f" -> Warning: Could not load {what.lower()} state: {e}. {what} will start from scratch.")
        # the reference replaces the caller's history object on resume (checkpoints never contain one, :530-536)
        try:
            agg = ckpt.get("history")
            agg.to(device)
            print(" -> Metrics History loaded.")
        except Exception:
            print(" -> No metric history saved")
            agg = MetricsHistory(num_classes, ignore_index)
        start_epoch = ckpt.get("epoch", 0)
        for k in best:
            best[k] = ckpt.get(k, best[k])
        print(f" -> Resuming training from epoch {start_epoch + 1}")
        print(f" -> Loaded best metrics: Dice={best['best_dev_dice']:.6f}, mIoU={best['best_dev_miou']:.6f}, "
              f"Loss={best['best_dev_loss']:.6f}")
        print(f" -> Notes from checkpoint: {ckpt.get('notes', 'N/A')}")
    else:
        print(f"Checkpoint file not found at {path}. Starting training from scratch.")

    print("\nStarting Training...")
    for t in range(start_epoch, epochs):
        print(f"Epoch {t + 1}\n-------------------------------")
        train_fn, eval_fn = (train_loop_prompt, eval_loop_prompt) if _prompt else (train_loop, eval_loop)
        train_fn(train_dataloader, model, train_loss_fn, optimizer, accumulation_steps, device, scheduler, target_size)
        val_loss, val_dice, val_miou = eval_fn(val_dataloader, model, val_loss_fn, device, target_size, agg)
        if save:
            torch.save({"epoch": t + 1, "history": agg}, os.path.join(model_save_dir, "metrics", model_save_name))
        if val_miou > best["best_dev_miou"]:
            best.update(best_dev_dice=val_dice, best_dev_miou=val_miou, best_dev_loss=val_loss)
            if save:
                print(f"Validation IoU score improved ({val_miou:.6f}). Saving model...")
                ckpt = {"epoch": t + 1, "model_state_dict": model.state_dict(),
                        "optimizer_state_dict": optimizer.state_dict(), **best,
                        "notes": f"Model saved based on best Micro Dice. Ignored index for metric: {ignore_index}"}
                if scheduler:
                    ckpt["scheduler_state_dict"] = scheduler.state_dict()
                if _prompt:
                    ckpt["history"] = agg                  # start_prompt keeps the history in the checkpoint, no MO_ copy
                torch.save(ckpt, path)
                if not _prompt:
                    torch.save({"epoch": t + 1, "model_state_dict": model.state_dict()},
                               os.path.join(model_save_dir, f"MO_{model_save_name}"))
        else:
            print(f"Validation IoU score did not improve from {best['best_dev_miou']:.6f}")
    print("\n--- Training Finished! ---")
    print(f"Best validation IoU score achieved: {best['best_dev_miou']:.6f}")
    print(f"Corresponding validation dice: {best['best_dev_dice']:.6f}")
    print(f"Corresponding validation loss: {best['best_dev_loss']:.6f}")
    print(f"Best model saved to: {os.path.join(model_save_dir, model_save_name)}")
    return best


def start_prompt(model_save_dir, model_save_name, model, optimizer, train_dataloader, val_dataloader, accumulation_steps,
                 device, train_loss_fn, val_loss_fn, target_size, scheduler=None, agg=None, load=True, save=True,
                 num_classes=4, ignore_index=3, epochs=100):
    """``start`` for prompt-based models (reference utils/training.py:299-450): batches are (image, heat-map, label), the
    loops are ``train_loop_prompt`` / ``eval_loop_prompt``, the checkpoint carries the metrics history and is loaded with
    pickle enabled, and no weights-only ``MO_`` copy is written."""
    return start(model_save_dir, model_save_name, model, optimizer, train_dataloader, val_dataloader, accumulation_steps,
                 device, train_loss_fn, val_loss_fn, target_size, scheduler, agg, load, save, num_classes, ignore_index, epochs,
                 _prompt=True)
