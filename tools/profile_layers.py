"""Per-launch CUDA-event table of one training step (run on the GPU box):
    python tools/profile_layers.py [--batch 64] > gpurun_out/layers.txt
Lists every C-ABI call with its time, algorithmic TFLOP/s (contractions) and share of the step."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from image_segmentation_b200 import _lib as L  # noqa: E402
from image_segmentation_b200.unet.unet import unet  # noqa: E402
from image_segmentation_b200.utils.synthetic import make_batch  # noqa: E402
from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--workload", default="unet", help="unet | autoencoder_recon | autoencoder_seg | clip | prompt (bench.py workloads)")
args = ap.parse_args()
dev = torch.device("cuda")
torch.manual_seed(0)
if args.workload == "unet":
    m = unet(3, 3).to(dev).train()
    opt = torch.optim.AdamW(m.parameters(), weight_decay=0.01)
    fn = WeightedDiceCELoss(smooth_dice=1, class_weights=torch.tensor([0.2, 1.0, 1.2]))
    x, y = make_batch(args.batch, 256, 256, 3, 3)
    x, y = x.to(dev), y.squeeze(1).to(dev)

    def step():
        loss = fn(m(x), y)
        loss.backward()
        opt.step()
        opt.zero_grad()
else:
    import bench
    m, inputs, fwd, _ = bench.build_family(args.workload, args.batch, dev)
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], weight_decay=0.01)

    def step():
        loss = fwd(m, *inputs)
        loss.backward()
        opt.step()
        opt.zero_grad()


for _ in range(3):
    step()
rec = []
torch.cuda.synchronize()
step_ms = 0.0
for _ in range(args.steps):
    # a spin kernel in front of every instrumented step lets the host enqueue the whole step before the GPU starts it:
    # the kernels then run back to back (sustained clocks), as inside the CUDA-graph step
    L.PROFILE_HOOK = rec
    torch.cuda._sleep(int(0.12 * 1.9e9))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    L.PROFILE_HOOK = None
    torch.cuda.synchronize()
    step_ms += e0.elapsed_time(e1) / args.steps
per = len(rec) // args.steps
rows = {}
for i, (kind, flops, a, b, label, _tag) in enumerate(rec):
    key = (i % per, kind, label)
    t = a.elapsed_time(b)
    r = rows.setdefault(key, [0.0, flops])
    r[0] += t / args.steps
print(f"step {step_ms:.3f} ms (instrumented), batch {args.batch}")
print(f"{'#':>4} {'kind':<14} {'layer':<18} {'ms':>8} {'TFLOP/s':>9} {'GFLOP':>10}")
tot = {}
for (idx, kind, label), (ms, flops) in sorted(rows.items()):
    tf = flops / (ms * 1e-3) / 1e12 if flops else 0.0
    print(f"{idx:>4} {kind:<14} {label:<18} {ms:>8.3f} {tf:>9.1f} {flops / 1e9:>10.1f}")
    tot.setdefault(kind, [0.0, 0.0])
    tot[kind][0] += ms
    tot[kind][1] += flops
print("---- by kind")
for k, (ms, flops) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:<14} {ms:>8.3f} ms  {100 * ms / step_ms:>5.1f}%  {flops / (ms * 1e-3) / 1e12 if flops else 0:>8.1f} TFLOP/s")
