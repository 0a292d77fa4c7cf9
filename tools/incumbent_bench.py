"""The GPU incumbent: the UNMODIFIED reference model / loss (staged under oracle/_ref by oracle/build_ref.py) run by
stock PyTorch on the same B200, same workload as bench.py (unet(3,3), batch 64, 256x256, AdamW, Dice+CE with class weights).

    python tools/incumbent_bench.py [--batch 64] [--steps 10] [--warmup 3] > gpurun_out/incumbent.jsonl

Prints one JSON line per variant:
  fp32_tf32          eager fp32 parameters/activations; cuDNN convolutions may use TF32 (torch's default), NCHW
  fp32_ieee          the same with torch.backends.cudnn.allow_tf32 = False (true fp32 -- what the CPU reference computes)
  bf16_autocast_cl   torch.autocast(bfloat16) + channels_last (the fastest stock configuration of the same model)
A step = forward + loss + backward + AdamW(fused) + zero_grad; the reference's train_loop (utils/training.py:38-60)
computes no metrics while training, so none are timed here (bench.py's own step DOES include MetricsHistory.accumulate).
This is test / measurement infrastructure: it imports oracle/, never the product.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import ref_shim  # noqa: E402

CLASS_W3 = [0.2046795970925636, 1.0271954434416883, 1.2293222812780409]
FLOP_PER_IMAGE_TRAIN = 288.8282e9


def make_batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, 256, 256, generator=g)
    y = torch.randint(0, 3, (n, 1, 256, 256), generator=g)
    return x, y


def run(variant, ref, batch, steps, warmup):
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = variant != "fp32_ieee"
    torch.backends.cuda.matmul.allow_tf32 = variant != "fp32_ieee"
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = ref.unet(3, 3).to(dev).train()
    autocast = variant == "bf16_autocast_cl"
    if autocast:
        model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.AdamW(model.parameters(), weight_decay=0.01, fused=True)
    loss_fn = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=torch.tensor(CLASS_W3).to(dev))
    x, y = make_batch(batch, 1234)
    x, y = x.to(dev), y.to(dev)
    if autocast:
        x = x.contiguous(memory_format=torch.channels_last)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            pred = model(x)
        loss = loss_fn(pred.float(), y.squeeze(1))
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"impl": "incumbent", "variant": variant, "metric": "unet256_train_images_per_sec", "value": batch / (ms * 1e-3),
            "unit": "images/s", "ms_per_step": ms, "batch": batch, "steps": steps, "warmup": warmup,
            "conv_tflops": batch * FLOP_PER_IMAGE_TRAIN / (ms * 1e-3) / 1e12, "loss": float(loss),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "torch": torch.__version__,
            "cudnn": torch.backends.cudnn.version(), "gpu": torch.cuda.get_device_name(0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--variants", default="fp32_tf32,fp32_ieee,bf16_autocast_cl")
    a = ap.parse_args()
    if not ref_shim.available():
        print(json.dumps({"impl": "incumbent", "unavailable": "reference sources not staged (run python -m oracle.build_ref "
                                                              "where /root/reference exists)"}))
        return
    ref = ref_shim.load()
    for v in a.variants.split(","):
        try:
            print(json.dumps(run(v, ref, a.batch, a.steps, a.warmup)), flush=True)
        except Exception as e:  # e.g. out of memory for one variant: report and go on
            print(json.dumps({"impl": "incumbent", "variant": v, "error": repr(e)[:300]}), flush=True)
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
