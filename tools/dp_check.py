"""2-rank GPU check of the data-parallel path (run with torchrun --nproc-per-node 2):
both ranks get the SAME shard, so the all-reduced (averaged) gradients must equal the single-GPU gradients; then ranks get
different shards and the result must equal the mean of the two independent single-GPU gradients."""
import datetime
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from image_segmentation_b200.parallel import DataParallelUNet  # noqa: E402
from image_segmentation_b200.unet.unet import unet  # noqa: E402
from image_segmentation_b200.utils.synthetic import make_batch  # noqa: E402
from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
fn = WeightedDiceCELoss(smooth_dice=1)


def grads(model, x, y):
    model.zero_grad()
    fn(model(x.to(dev)), y.squeeze(1).to(dev)).backward()
    return {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30)).item()


torch.manual_seed(0)
single = unet(3, 3)
single.precision = "fp32"
single = single.to(dev).train()
torch.manual_seed(0)
model = unet(3, 3)
model.precision = "fp32"
model = model.to(dev).train()
dp = DataParallelUNet(model, bucket_mb=8)
shards = [make_batch(2, 64, 64, 3, 3, seed=50 + r) for r in range(world)]
# (1) identical shards on every rank
g_dp = grads(model, *shards[0])
g_1 = grads(single, *shards[0])
worst = max(rel(g_dp[k], g_1[k]) for k in g_1 if g_1[k].abs().max() > 0)
print(f"[rank {rank}] identical shards: worst rel-L2 vs single GPU = {worst:.2e}")
assert worst < 1e-5
# (2) different shards: mean of the independent gradients
g_dp = grads(model, *shards[rank])
ref = None
for r in range(world):
    gr = grads(single, *shards[r])
    ref = gr if ref is None else {k: ref[k] + gr[k] for k in gr}
ref = {k: v / world for k, v in ref.items()}
worst = max(rel(g_dp[k], ref[k]) for k in ref if ref[k].abs().max() > 0)
print(f"[rank {rank}] different shards: worst rel-L2 vs mean of single-GPU gradients = {worst:.2e}")
assert worst < 1e-5
dist.barrier()
if rank == 0:
    print("dp_check OK")
dist.destroy_process_group()
