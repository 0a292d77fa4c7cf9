"""Throughput THROUGH the drop-in train_loop (utils/training.py:18-64) with a real DataLoader, for the reference's own
configuration (micro-batch 2 x 32 accumulation steps, unet/unet.ipynb:41-42,64) and for true batch 64:

    python tools/loop_bench.py > gpurun_out/loop_bench.jsonl

One JSON line per configuration: images/s over the second epoch (the first one contains the eager warm-up steps and the
CUDA-graph capture), with graph replay on and off."""
import contextlib
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.utils.data import DataLoader, TensorDataset  # noqa: E402

from image_segmentation_b200.unet.unet import unet  # noqa: E402
from image_segmentation_b200.utils.synthetic import make_batch  # noqa: E402
from image_segmentation_b200.utils.training import train_loop  # noqa: E402
from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss  # noqa: E402

dev = torch.device("cuda")
x, y = make_batch(512, 256, 256, 3, 3, seed=1)
ds = TensorDataset(x, y.to(torch.uint8))
for micro, accum in ((64, 1), (2, 32), (8, 8)):
    for graph in ("1", "0"):
        os.environ["UNETK_TRAIN_GRAPH"] = graph
        torch.manual_seed(0)
        m = unet(3, 3).to(dev)
        opt = torch.optim.AdamW(m.parameters(), weight_decay=0.01, capturable=True)
        fn = WeightedDiceCELoss(smooth_dice=1, class_weights=torch.tensor([0.2046795970925636, 1.0271954434416883, 1.2293222812780409]))
        loader = DataLoader(ds, batch_size=micro, shuffle=False, pin_memory=True)
        times = []
        for epoch in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                loss = train_loop(loader, m, fn, opt, accum, dev, None, 256)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        print(json.dumps({"micro_batch": micro, "accumulation_steps": accum, "graph": graph == "1",
                          "images_per_s": len(ds) / min(times[1:]), "epoch_s": [round(t, 3) for t in times],
                          "last_epoch_loss": loss, "images_per_epoch": len(ds)}), flush=True)
        del m, opt
        torch.cuda.empty_cache()
