#!/usr/bin/env python
"""bench.py -- U-Net 256x256 training throughput (images/s) on N B200s; one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (bf16 tcgen05 tier)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path on the host cores
    python bench.py --workload autoencoder_recon|autoencoder_seg|clip|prompt   # the other model families (one GPU)

Workload (BASELINE.json configs[1]): unet(3,3) bf16 training step, batch 64 per GPU at 256x256,
WeightedDiceCELoss(smooth_dice=1, class weights), AdamW(lr 1e-3, wd 0.01), MetricsHistory.accumulate,
synthetic U[0,1) images and i.i.d. 3-class labels, random-init weights.

A step = forward + loss + backward (+ gradient exchange for N > 1) + optimizer.step + zero_grad + metrics.
`value`  : images/s of the whole job with the batch already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through the public API with the batch in pinned HOST memory: H2D of X (fp32) and y (uint8)
           and a D2H read of the loss inside the timed region every step.
`roofline`: tensor-pipe roofline of the tcgen05 contraction kernels: algorithmic FLOPs / CUDA-event time of those
           launches in three instrumented eager steps, each enqueued behind a spin kernel so that the launches run back
           to back (refused when they cover less than half of the step); plus the whole-step figures.
`cpu_baseline`: the UNMODIFIED reference modules staged under oracle/_ref (kind "reference"; the oracle port, kind
           "port", only if they did not travel) timed on the host cores for a bounded sample (batch 4 steps).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "unet256_train_images_per_sec"
UNIT = "images/s"
CLASS_W3 = [0.2046795970925636, 1.0271954434416883, 1.2293222812780409]
FLOP_PER_IMAGE_TRAIN = 288.8282e9   # BASELINE.md section 2 (conv/convT/head, 2 FLOP per MAC, fwd + dgrad + wgrad)
H = W = 256


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle port (the only place besides tests/ and smoke() that executes oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, batch=4):
    """Config[0] of BASELINE.json on the host cores.  Uses the UNMODIFIED reference modules staged under oracle/_ref/
    (oracle/build_ref.py; kind "reference") and falls back to the oracle port (kind "port") when they did not travel."""
    import torch
    from image_segmentation_b200.utils.synthetic import make_batch
    from oracle import ref_shim
    # every host core: torchrun exports OMP_NUM_THREADS=1 for its workers, which would time the CPU path on one thread
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    torch.set_num_threads(max(1, cores))
    torch.manual_seed(0)
    w = torch.tensor(CLASS_W3)
    x, y = make_batch(batch, H, W, 3, 3, seed=1234)
    if ref_shim.available():
        ref = ref_shim.load()
        kind = "reference"
        m = ref.unet(3, 3).train()                                   # unet/unet.py:67
        loss_fn = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=w)   # utils/weighted_loss.py:102
        agg = ref.MetricsHistory(3)                                  # utils/MetricsHistory.py:9

        def loss_of(pred):
            return loss_fn(pred, y.squeeze(1))

        def metrics_of(pred):
            for j in range(batch):
                agg.accumulate(pred[j].detach(), y[j])
    else:
        from oracle import loss_oracle, metrics_oracle, unet_oracle
        kind = "port"
        m = unet_oracle.OracleUNet(3, 3, native_ops=True).train()   # the reference's own ATen operators (F.batch_norm, ...)

        def loss_of(pred):
            return loss_oracle.dice_ce_loss(pred, y, smooth_dice=1.0, class_weights=w, dtype=torch.float32)

        def metrics_of(pred):
            for j in range(batch):
                metrics_oracle.confusion_counts(pred[j].detach().numpy(), y[j].numpy(), 3)
    opt = torch.optim.AdamW(m.parameters(), weight_decay=0.01)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        pred = m(x)
        loss = loss_of(pred)
        loss.backward()
        opt.step()
        opt.zero_grad()
        metrics_of(pred)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return batch, times, torch.get_num_threads(), kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch, times, threads, kind = cpu_reference_steps(args.steps, args.warmup)
    total = sum(times)
    value = batch * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "unet(3,3) 256x256 training step (fwd+loss+bwd+AdamW+metrics), CPU fp32, "
                               f"bounded sample: batch {batch} per step instead of 64"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{len(times)} steps of batch {batch} at 256x256 on {threads} threads "
                                   f"(os.cpu_count()={os.cpu_count()})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from image_segmentation_b200 import _lib as L
    from image_segmentation_b200.parallel import DataParallelUNet
    from image_segmentation_b200.unet.unet import unet
    from image_segmentation_b200.utils.MetricsHistory import MetricsHistory
    from image_segmentation_b200.utils.synthetic import make_batch
    from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    L.lib()  # fail loudly now if the extension is missing
    peaks = load_peaks()
    B = args.batch

    torch.manual_seed(0)
    model = unet(3, 3)
    model.precision = "bf16"
    model = model.to(dev).train()
    # UNETK_DP_BUCKET_MB / UNETK_DP_COMPRESS: scaling experiments (defaults: 25 MB buckets, fp32 gradients on the wire)
    dp_kw = dict(bucket_mb=float(os.environ.get("UNETK_DP_BUCKET_MB", "25")), compress=os.environ.get("UNETK_DP_COMPRESS") or None,
                 exchange=os.environ.get("UNETK_DP_EXCHANGE", "auto"))
    # UNETK_DP_DISABLE=1: N independent replicas (no gradient exchange) -- isolates the straggler / shared-power effect of
    # running N GPUs of one box at once from the cost of the all-reduce
    dp_off = os.environ.get("UNETK_DP_DISABLE", "0") == "1"
    dp = DataParallelUNet(model, **dp_kw) if (world > 1 and not dp_off) else None
    # same update rule as the reference's AdamW; fused = one kernel, capturable = usable inside a CUDA graph
    opt = torch.optim.AdamW(model.parameters(), weight_decay=0.01, fused=True, capturable=True)
    loss_fn = WeightedDiceCELoss(smooth_dice=1, class_weights=torch.tensor(CLASS_W3))
    agg = MetricsHistory(3)
    x_cpu, y_cpu = make_batch(B, H, W, 3, 3, seed=1234 + rank)
    x_pin = x_cpu.pin_memory()
    y_pin = y_cpu.to(torch.uint8).pin_memory()          # the reference's dataset yields uint8 label maps
    x_dev, y_dev = x_cpu.to(dev), y_cpu.squeeze(1).to(dev)

    def step_resident():
        pred = model(x_dev)
        loss = loss_fn(pred, y_dev)
        loss.backward()
        opt.step()
        opt.zero_grad()
        agg.accumulate(pred.detach(), y_dev)
        return loss

    from image_segmentation_b200.utils.prefetch import AsyncScalarReader, DevicePrefetcher

    def run_e2e(steps):
        # exactly the body of the drop-in train_loop (image_segmentation_b200/utils/training.py; reference
        # utils/training.py:45-60) + metrics: every step copies ITS batch from pinned host memory (the copy of batch i+1
        # is in flight while step i computes, as train_loop does it) and reads the loss back to the host (every step's
        # loss, one step late -- AsyncScalarReader -- so the launch queue never drains; the last one before returning)
        # With one GPU the step itself is the package's GraphedTrainStep (what train_loop replays for a capturable
        # optimizer); data-parallel runs use the eager step.
        out = 0.0
        reader = AsyncScalarReader(dev)
        for X, y in DevicePrefetcher(((x_pin, y_pin) for _ in range(steps)), dev):
            if graphed is not None:
                loss = graphed(X, y)
            else:
                pred = model(X)
                loss = loss_fn(pred, y.squeeze(1))
                loss.backward()
                opt.step()
                opt.zero_grad()
                agg.accumulate(pred.detach(), y.squeeze(1))
            reader.push(loss)                            # D2H read of the step's result
            for v in reader.ready():
                out += v
        for v in reader.drain():
            out += v
        return out

    graphed = None
    rank_ms = []          # per-rank ms/step of the last timed() call (world > 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            allms = [torch.zeros(1, device=dev) for _ in range(world)]
            dist.all_gather(allms, torch.tensor([ms], device=dev))
            rank_ms[:] = [float(t.item()) / max(steps, 1) for t in allms]
            ms = max(float(t.item()) for t in allms)
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # the whole step replays as ONE CUDA graph (image_segmentation_b200.utils.graph)
    graphed = None
    # (data parallel: the bucketed NCCL all-reduces are captured with the step; UNETK_DP_GRAPH=0 keeps N > 1 eager)
    if not args.no_graph and (world == 1 or os.environ.get("UNETK_DP_GRAPH", "1") == "1"):
        from image_segmentation_b200.utils.graph import GraphedTrainStep
        try:
            L.COUNTERS["launches"] = 0
            graphed = GraphedTrainStep(model, loss_fn, opt, x_dev, y_dev, metrics=agg, warmup=1)
            launches_per_step = L.COUNTERS["launches"] // 2      # 1 warm-up step + 1 captured step
        except Exception as e:   # report and measure the eager path instead
            print(f"[bench] CUDA graph capture failed, timing the eager path: {e!r}", file=sys.stderr)
            graphed = None
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if graphed is not None:
        for _ in range(3):
            graphed(x_dev, y_dev)
        ms = timed(lambda: graphed(x_dev, y_dev), args.steps)
        launches = launches_per_step * args.steps
    else:
        L.COUNTERS["launches"] = 0
        ms = timed(step_resident, args.steps)
        launches = L.COUNTERS["launches"]
    clocks = sampler.stop() if rank == 0 else None
    rank_ms_value = list(rank_ms)
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)

    run_e2e(2)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # ---- per-launch instrumentation of the contraction kernels (extra steps, not part of `value`; every rank runs
    #      them -- they contain collectives -- but only rank 0 records events).
    #      The instrumented steps are EAGER (one CUDA event pair per launch), so the host is slower than the GPU.  To time
    #      the kernels the way they run inside the graphed step -- back to back, at sustained clocks -- every
    #      instrumented step is enqueued behind a spin kernel (torch.cuda._sleep) long enough for the host to finish
    #      enqueueing the whole step; `share_of_step` (contraction time / GPU time of the step, spin excluded) shows
    #      whether that worked, and no roofline figure is printed when it is below 0.5. ----
    roof = None
    prof_steps = 3
    spin_ms_used = args.spin_ms

    def instrumented_steps(spin_ms):
        """One discarded step (creates the events and warms the pools) + prof_steps recorded ones.  Returns the records
        and, per step, (GPU ms of the step without the spin, contraction-kernel ms)."""
        cycles = int(spin_ms * 1e-3 * 1.9e9)
        recs, per_step = [], []
        torch.cuda.synchronize()
        for i in range(prof_steps + 1):
            mine = []
            if rank == 0:
                L.PROFILE_HOOK = mine
            if cycles > 0:
                torch.cuda._sleep(cycles)
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            step_resident()
            m1.record()
            L.PROFILE_HOOK = None
            torch.cuda.synchronize()
            if i == 0:
                continue
            recs.extend(mine)
            per_step.append((m0.elapsed_time(m1),
                             sum(r[2].elapsed_time(r[3]) for r in mine if r[0] in ("conv", "wgrad"))))
        return recs, per_step

    # a host hiccup (a slow first enqueue on a fresh box) leaves gaps between the launches of a step: the kernel durations
    # stay valid, the step's share does not.  Retry with a longer spin (every rank takes the same decision).
    for attempt in range(3):
        records, per_step = instrumented_steps(spin_ms_used)
        best_share = max((k / t if t > 0 else 0.0) for t, k in per_step) if rank == 0 else 1.0
        again = torch.tensor([1 if best_share < 0.5 else 0], device=dev, dtype=torch.int32)
        if world > 1:
            dist.broadcast(again, 0)
        if int(again.item()) == 0 or args.spin_ms <= 0:
            break
        spin_ms_used *= 2
    if rank == 0:
        # the step whose launches ran back to back best stands for the step time; every recorded launch counts for the rate
        step_ms, _ = min(per_step, key=lambda tk: tk[0])
        fam = {}

        def family(r):
            if r[0] == "wgrad":
                return "wgrad"
            return "conv3x3" if r[5] == L.MODE_3X3 else ("convT" if r[5] in (L.MODE_CONVT, L.MODE_CONVT_GATHER) else "conv1x1")
        for r in records:
            if r[0] in ("conv", "wgrad"):
                f = fam.setdefault(family(r), [0.0, 0.0, 0])
                f[0] += r[1]
                f[1] += r[2].elapsed_time(r[3])
                f[2] += 1
        flops = sum(f[0] for f in fam.values())
        kms = sum(f[1] for f in fam.values())
        by_kind = {}
        for r in records:
            by_kind.setdefault(r[0], [0.0, 0.0])
            by_kind[r[0]][0] += r[2].elapsed_time(r[3]) / prof_steps
            by_kind[r[0]][1] += r[1] / prof_steps
        share = max((k / t if t > 0 else 0.0) for t, k in per_step)
        achieved = flops / (kms * 1e-3) / 1e12 if kms > 0 else 0.0
        valid = share >= 0.5
        dominant = None
        dpath = os.path.join(ROOT, "profiles", "dominant_launch.json")
        if os.path.isfile(dpath):
            dominant = json.load(open(dpath))
        step_tflops = (B * FLOP_PER_IMAGE_TRAIN / (ms_per_step * 1e-3)) / 1e12
        roof = {"bound": "tensor", "achieved": achieved if valid else None, "peak": peaks["sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["sustained"] if valid else None,
                "frac_of_burst": achieved / peaks["burst"] if valid else None,
                "traffic": dominant["traffic_bytes"] if dominant else None,
                "peak_source": peaks["source"] + ", sustained figure (kernels timed back to back inside a full step); "
                               "frac_of_burst uses the burst figure",
                "kernel": "tc::tc_conv_halo2_kernel / tc_conv2_kernel / tc_wgrad3x3*_kernel (tcgen05 implicit-GEMM family)",
                "how": f"sum of algorithmic FLOPs / sum of CUDA-event durations over every unetk_conv and unetk_wgrad launch of "
                       f"{prof_steps} instrumented eager steps, each enqueued behind a {spin_ms_used:.0f} ms spin kernel so that "
                       "the launches run back to back (share_of_step / instrumented_step_ms: the best of those steps; "
                       "the spin is doubled and the steps repeated when a host stall left gaps); refused (null) when "
                       "share_of_step < 0.5; `traffic` = ncu DRAM bytes of "
                       "the dominant launch (profiles/dominant_launch.json)",
                "valid": valid,
                "share_of_step": share,
                "instrumented_step_ms": step_ms,
                "instrumented_steps_ms": [round(t, 3) for t, _ in per_step],
                "families": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else None, "ms_per_step": v[1] / prof_steps,
                                 "launches_per_step": v[2] // prof_steps} for k, v in sorted(fam.items())},
                "dominant_launch": dominant,
                "ms_per_step_by_kernel": {k: round(v[0], 3) for k, v in sorted(by_kind.items())},
                "step_tflops": step_tflops,
                "step_frac": step_tflops / peaks["sustained"],
                "step_frac_of_burst": step_tflops / peaks["burst"]}

    if dp is not None and dp._trace is not None:
        # UNETK_DP_TRACE=1: ten eager steps back to back with CUDA events around the pieces of the gradient exchange
        dp._trace.clear()
        for _ in range(10):
            step_resident()
        torch.cuda.synchronize()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b, times, threads, kind = cpu_reference_steps(steps=3, warmup=1)
        v = b * len(times) / sum(times)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"3 steps of batch {b} at 256x256 (fp32, {threads} threads, os.cpu_count()={os.cpu_count()})"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"unet(3,3) 256x256 training step, batch {B}/GPU, bf16 activations + fp32 master weights, "
                                   "WeightedDiceCELoss + AdamW + MetricsHistory",
                       "global_batch": B * world, "parallelism": f"dp{world}",
                       **({"dp": dict(dp_kw, exchange="none (independent replicas)" if dp_off else
                                      ("NVLS multicast all-reduce kernel (unetk_nvls_allreduce_f32)" if (dp is not None and dp._nvls)
                                       else "bucketed NCCL all-reduce")),
                           "rank_ms_per_step": [round(v, 3) for v in rank_ms_value],
                           **({"exchange_trace_ms": dp.trace_summary()} if (dp is not None and dp._trace) else {})} if world > 1 else {}),
                       "launch": "one CUDA graph per step" if graphed is not None else "eager launches",
                       "e2e_path": "pinned host batch -> DevicePrefetcher (copy of batch i+1 overlaps step i) -> "
                                   + ("GraphedTrainStep" if graphed is not None else "eager step")
                                   + " -> AsyncScalarReader (every step's loss read back on the host)",
                       "l2": "per-step working set ~20 GB >> 126 MB L2 (no flush needed)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_pin.numel() * 4 + y_pin.numel(),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down without ncclCommDestroy: a live CUDA graph that captured collectives makes destroy_process_group()
        # hang (seen on 2 x B200).  Every rank has finished its work once the barrier returns; leave immediately.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------------------
# other model families on the same kernels (BASELINE.json configs[2..4]); single GPU
# ------------------------------------------------------------------------------------------------
def build_family(workload, B, dev, reference=False):
    """Returns (model, inputs on `dev` (list), step(*inputs) -> loss, description).  ``reference=True`` builds the
    UNMODIFIED reference modules (oracle/_ref) instead of this package's."""
    import torch
    CW4 = torch.tensor([0.2046795970925636, 1.0271954434416883, 1.2293222812780409, 1.5388026781877073])
    g = torch.Generator().manual_seed(1234)
    if reference:
        from oracle import ref_shim
        ref = ref_shim.load()
    torch.manual_seed(0)
    if workload in ("autoencoder_recon", "autoencoder_seg"):
        hw = 256
        x = torch.rand(B, 3, hw, hw, generator=g)
        if reference:
            ae = ref.autoencoder_mod
            Dice = ref.WeightedDiceCELoss
        else:
            from image_segmentation_b200.autoencoder import autoencoder as ae
            from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss as Dice
        if workload == "autoencoder_recon":
            model = ae.ReconstructionAutoencoder(3)
            mse = torch.nn.MSELoss()
            inputs = [x]
            fwd = lambda m, x: mse(m(x), x)  # noqa: E731
            desc = "ReconstructionAutoencoder(3) 256x256 reconstruction pre-training step (MSE)"
        else:
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                model = ae.SegmentationAutoencoder(3, 64, 4, freeze_encoder=True)
            y = torch.randint(0, 4, (B, hw, hw), generator=g)
            loss_fn = Dice(smooth_dice=1, class_weights=CW4.to(dev) if reference else CW4)
            inputs = [x, y]
            fwd = lambda m, x, y: loss_fn(m(x), y)  # noqa: E731
            desc = "SegmentationAutoencoder(3, 64, 4, freeze_encoder=True) 256x256 training step (Dice+CE)"
    elif workload in ("clip", "prompt"):
        from oracle import ref_shim
        hw = 224
        x = torch.rand(B, 3, hw, hw, generator=g)
        y = torch.randint(0, 4, (B, hw, hw), generator=g)
        with ref_shim.random_init_clip_vit(seed=0):
            if reference:
                clip = ref.clip_mod.ClipUNet()
            else:
                from image_segmentation_b200.clip.clipunet import ClipUNet
                clip = ClipUNet()
            if workload == "clip":
                model = clip
                Dice = ref.WeightedDiceCELoss if reference else __import__(
                    "image_segmentation_b200.utils.weighted_loss", fromlist=["WeightedDiceCELoss"]).WeightedDiceCELoss
                loss_fn = Dice(smooth_dice=1, class_weights=CW4.to(dev) if reference else CW4)
                inputs = [x, y]
                fwd = lambda m, x, y: loss_fn(m(x), y)  # noqa: E731
                desc = "ClipUNet (frozen random-init CLIP ViT-B/16 + U-Net decoder) 224x224 training step (Dice+CE)"
            else:
                heat = torch.rand(B, 1, hw, hw, generator=g)
                stable_log = lambda t: torch.log(t + 1e-9)  # noqa: E731
                if reference:
                    model = ref.prompt_mod.PromptModel()
                    NLL = ref.loss_mod.WeightedDiceNLLLoss
                else:
                    from image_segmentation_b200.prompt_based.prompt import PromptModel
                    from image_segmentation_b200.utils.weighted_loss import WeightedDiceNLLLoss as NLL
                    model = PromptModel(clip=clip)
                loss_fn = NLL(smooth_dice=1, class_weights=CW4.to(dev) if reference else CW4, apply_softmax=False, nll_nonlin=stable_log)
                inputs = [x, heat, y]
                fwd = lambda m, x, h, y: loss_fn(m(x, h), y)  # noqa: E731
                desc = "PromptModel (frozen ClipUNet + unet(4,1) + probability composition) 224x224 training step (Dice+NLL)"
    else:
        raise SystemExit(f"unknown workload {workload}")
    model = model.to(dev).train()
    if not reference and workload in ("clip", "prompt") and os.environ.get("UNETK_VIT_FP32", "0") != "1":
        # the frozen third-party ViT under bf16 autocast (this package's tier is bf16); UNETK_VIT_FP32=1 keeps the reference's fp32
        (model if workload == "clip" else model.clip).vit_autocast_dtype = torch.bfloat16
        desc += "; frozen ViT under torch.autocast(bf16)"
    return model, [t.to(dev) for t in inputs], fwd, desc


def run_family_arm(args):
    import torch
    from image_segmentation_b200 import _lib as L
    from image_segmentation_b200.utils.graph import GraphedStep
    from image_segmentation_b200.utils.prefetch import AsyncScalarReader
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("--workload other than unet is a single-GPU measurement")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L.lib()
    peaks = load_peaks()
    B = args.batch
    model, inputs, fwd, desc = build_family(args.workload, B, dev)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.01, fused=True, capturable=True)

    def step(*inp):
        loss = fwd(model, *inp)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.detach()

    for _ in range(max(args.warmup, 3)):
        step(*inputs)
    graphed = None
    if not args.no_graph:
        try:
            graphed = GraphedStep(step, inputs, models=[m for m in model.modules() if hasattr(m, "_engine")], optimizer=opt, warmup=1)
        except Exception as e:
            print(f"[bench] CUDA graph capture failed, timing the eager path: {e!r}", file=sys.stderr)
    run = (lambda: graphed(*inputs)) if graphed is not None else (lambda: step(*inputs))
    for _ in range(3):
        run()
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.COUNTERS["launches"] = 0
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    value = B * args.steps / (ms / 1e3)
    # end to end: the step's inputs come from pinned host memory every step, the loss is read back on the host
    pinned = [t.cpu().pin_memory() for t in inputs]
    slots = [[torch.empty_like(t) for t in inputs] for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)

    def run_e2e(steps):
        reader, total, evs = AsyncScalarReader(dev), 0.0, [None, None]
        for i in range(steps):
            k = i % 2
            with torch.cuda.stream(copy_stream):
                if evs[k] is not None:
                    copy_stream.wait_event(evs[k])
                for d, s in zip(slots[k], pinned):
                    d.copy_(s, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            torch.cuda.current_stream().wait_event(ready)
            loss = graphed(*slots[k]) if graphed is not None else step(*slots[k])
            evs[k] = torch.cuda.Event()
            evs[k].record()
            reader.push(loss)
            total += sum(reader.ready())
        return total + sum(reader.drain())
    run_e2e(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    # contraction roofline from instrumented eager steps behind a spin kernel (see run_gpu_arm)
    records, marks = [], []
    for i in range(4):                                  # the first instrumented step is discarded (creates the events)
        L.PROFILE_HOOK = records if i else []
        torch.cuda._sleep(int(args.spin_ms * 1e-3 * 1.9e9))
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        step(*inputs)
        m1.record()
        L.PROFILE_HOOK = None
        torch.cuda.synchronize()
        if i:
            marks.append((m0, m1))
    step_ms = sum(a.elapsed_time(b) for a, b in marks) / 3
    conv = [r for r in records if r[0] in ("conv", "wgrad")]
    flops = sum(r[1] for r in conv)
    kms = sum(r[2].elapsed_time(r[3]) for r in conv)
    by_kind = {}
    for r in records:
        by_kind[r[0]] = by_kind.get(r[0], 0.0) + r[2].elapsed_time(r[3]) / 3
    share = kms / 3 / step_ms
    achieved = flops / (kms * 1e-3) / 1e12 if kms else 0.0
    flop_per_image = flops / 3 / B
    roof = {"bound": "tensor", "achieved": achieved if share >= 0.3 else None, "peak": peaks["sustained"], "unit": "TFLOP/s",
            "frac": achieved / peaks["sustained"] if share >= 0.3 else None, "frac_of_burst": achieved / peaks["burst"],
            "traffic": None, "share_of_step": share, "instrumented_step_ms": step_ms,
            "ms_per_step_by_kernel": {k: round(v, 3) for k, v in sorted(by_kind.items())},
            "algorithmic_gflop_per_image": flop_per_image / 1e9,
            "step_tflops": B * flop_per_image / (ms / args.steps * 1e-3) / 1e12,
            "note": "contraction kernels of this package only; time spent in torch's frozen ViT (clip / prompt workloads) is part "
                    "of the step but not of `achieved`"}
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = family_cpu_baseline(args.workload)
        except Exception as e:
            cpu = {"unavailable": repr(e)[:200]}
    line = {"metric": f"{args.workload}_train_images_per_sec", "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc + f", batch {B}, bf16 activations + fp32 master weights, AdamW",
                       "launch": "one CUDA graph per step" if graphed is not None else "eager launches"},
            "clocks": clocks,
            "e2e": {"value": B * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in pinned),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": L.COUNTERS["launches"] if graphed is None else None,
            "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def family_cpu_baseline(workload, batch=2, steps=2):
    """The UNMODIFIED reference model of the same family on the host cores (bounded sample: batch 2)."""
    import torch
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    torch.set_num_threads(max(1, cores))
    model, inputs, fwd, _ = build_family(workload, batch, torch.device("cpu"), reference=True)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.01)
    times = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        loss = fwd(model, *inputs)
        loss.backward()
        opt.step()
        opt.zero_grad()
        if i > 0:
            times.append(time.perf_counter() - t0)
    return {"value": batch * len(times) / sum(times), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"{len(times)} steps of batch {batch} ({torch.get_num_threads()} threads, fp32)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="unet", choices=["unet", "autoencoder_recon", "autoencoder_seg", "clip", "prompt"],
                    help="unet = the headline (BASELINE.json configs[1]); the others are the model families of configs[2..4]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the CUDA-graph step")
    ap.add_argument("--spin-ms", type=float, default=120.0,
                    help="length of the spin kernel in front of each instrumented (roofline) step; 0 disables it")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload != "unet":
        run_family_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
