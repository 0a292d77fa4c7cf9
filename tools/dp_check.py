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
dp = DataParallelUNet(model, bucket_mb=8, exchange=os.environ.get("UNETK_DP_EXCHANGE", "auto"))
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
# (3) the graphed data-parallel step (bench.py at N > 1): all-reduces captured inside the CUDA graph.  Every rank trains
# on a DIFFERENT shard; if the captured collectives ran, the replicas stay bit-identical and their update equals the
# eager data-parallel update from the same state.
from image_segmentation_b200.utils.graph import GraphedTrainStep  # noqa: E402

state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
# plain SGD: the update is linear in the gradient, so atomics-level noise is not amplified the way Adam's g/|g| does it
opt = torch.optim.SGD(model.parameters(), lr=1e-3)
xs, ys = shards[rank][0].to(dev), shards[rank][1].squeeze(1).to(dev)
flat0 = torch.cat([state0[k].flatten() for k, _ in model.named_parameters()])


def cur_flat():
    torch.cuda.synchronize()
    return torch.cat([p.detach().flatten() for p in model.parameters()])


step = GraphedTrainStep(model, fn, opt, xs, ys, warmup=1)         # 1 eager + 0 replayed steps so far
trace_g = [(cur_flat() - flat0).norm().item()]
for _ in range(3):
    step(xs, ys)
    trace_g.append((cur_flat() - flat0).norm().item())
flat = cur_flat()
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
for r in range(world):
    assert torch.equal(gathered[r], gathered[0]), "replicas diverged: captured all-reduce did not run"
# eager reference: same initial state, 4 optimiser steps (1 warm-up + 3 replays) with the eager data-parallel backward
model.load_state_dict(state0)
opt2 = torch.optim.SGD(model.parameters(), lr=1e-3)
trace_e = []
for _ in range(4):
    opt2.zero_grad()
    fn(model(xs), ys).backward()
    opt2.step()
    trace_e.append((cur_flat() - flat0).norm().item())
flat2 = cur_flat()
print(f"[rank {rank}] |update| graph path {['%.3e' % v for v in trace_g]}  eager path {['%.3e' % v for v in trace_e]}")
err = rel(flat - flat0, flat2 - flat0)        # compare the UPDATES: a missing / unaveraged bucket would show as O(1)
print(f"[rank {rank}] graphed DP vs eager DP after 4 SGD steps: rel-L2 of the parameter update = {err:.2e}; "
      f"replicas bit-identical")
assert err < 5e-2                             # 4 steps of chaotic dynamics (ReLU / max-pool flips) on top of 3e-7 gradient noise
dist.barrier()
torch.cuda.synchronize()
if rank == 0:
    print("dp_check OK; exchange:", "nvls kernel" if dp._nvls else "nccl")
sys.stdout.flush()
os._exit(0)      # no destroy_process_group() while a graph with captured collectives is alive (it hangs)
