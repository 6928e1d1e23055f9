"""Kernel timeline of one RoadMapBCE training step (torch.profiler / CUPTI): start, duration and stream of every kernel,
to check what overlaps what.  Works with world size 1 or under torchrun; rank 0 prints."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
from driving_dirty_b200.optim import FusedAdam
from driving_dirty_b200.synthetic import scene_batch
model = bench.build_model("bf16", dev)
params = [p for p in model.parameters() if p.requires_grad]
mc = os.environ.get("MC")
opt = FusedAdam(params, lr=1e-3, overlap_backward=os.environ.get("OVERLAP", "1") == "1",
                multicast=None if mc is None else mc == "1")
views, road = scene_batch(32, 256, 306, seed=1 + rank)
views, road = views.to(dev), road.to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    out = model.training_step((views, None, road), 1)
    out["loss"].backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    half = len(evs) // 2
    print(f"{len(evs)} device events in 2 steps; second step:")
    for e in evs[half:]:
        print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  {e.time_range.elapsed_us() / 1e3:7.3f} ms  {e.name[:90]}")
if world > 1: dist.destroy_process_group()
