"""Per-step device times of the RoadMapBCE training step (CUDA events around every step), rank 0 prints."""
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
views, road = scene_batch(32, 256, 306, seed=1 + rank)
views, road = views.to(dev), road.to(dev)
n = int(os.environ.get("STEPS", 24))
lines = []
for mode in os.environ.get("MODES", "auto").split():
    model = bench.build_model("bf16", dev)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = FusedAdam(params, lr=1e-3, overlap_backward=os.environ.get("OVERLAP", "1") == "1",
                    multicast=None if mode == "auto" else bool(int(mode)))
    def step():
        opt.zero_grad(set_to_none=True)
        out = model.training_step((views, None, road), 1)
        out["loss"].backward()
        opt.step()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    evs[0].record()
    for i in range(n):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    t = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    lines.append(f"world {world} multicast={opt.uses_multicast}: " + " ".join(f"{x:.2f}" for x in t))
    del model, params, opt
    torch.cuda.empty_cache()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/step_times_n{world}.txt", "w") as f:
        f.write("\n".join(lines) + "\n")
if world > 1: dist.destroy_process_group()
