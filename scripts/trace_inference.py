"""Per-kernel device time of one ModelLoader.get_binary_road_map call (torch.profiler / CUPTI), summed by kernel name.
BATCH (default 256) scenes of raw camera bytes, device resident; DTYPE bf16 (default) or fp32."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from driving_dirty_b200.model_loader import ModelLoader
from driving_dirty_b200.synthetic import random_roadmap_model, scene_batch_bytes
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
b = int(os.environ.get("BATCH", "256"))
dtype = os.environ.get("DTYPE", "bf16")
model = random_roadmap_model(bench.HIDDEN, bench.LATENT, bench.VIEW_H, bench.VIEW_W, dtype=dtype, device=dev)
loader = ModelLoader(model, device=dev)
views, _ = scene_batch_bytes(b, bench.VIEW_H, bench.VIEW_W, seed=3)
views = views.to(dev)
for _ in range(3):
    loader.get_binary_road_map(views, as_bytes=True)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    loader.get_binary_road_map(views, as_bytes=True)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tot = collections.OrderedDict()
for e in evs:
    k = e.name[:100]
    n, t = tot.get(k, (0, 0.0))
    tot[k] = (n + 1, t + e.time_range.elapsed_us() / 1e3)
span = (evs[-1].time_range.end - evs[0].time_range.start) / 1e3
busy = sum(t for _, t in tot.values())
print(f"batch {b} {dtype}: {len(evs)} device events, span {span:.3f} ms, sum of kernel times {busy:.3f} ms")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:8.3f} ms  {100 * t / span:5.1f} %  x{n:<4d} {k}")

if os.environ.get("GAPS"):
    print("gaps > 20 us between consecutive device events:")
    for a, b in zip(evs, evs[1:]):
        gap = (b.time_range.start - a.time_range.end) / 1e3
        if gap > 0.02:
            print(f"{gap:8.3f} ms after {a.name[:60]}  ->  {b.name[:60]}")
    cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.time_range.elapsed_us() > 100]
    cpu.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    print("host-side ops longer than 100 us (start relative to the first kernel, duration):")
    for e in cpu:
        print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  {e.time_range.elapsed_us() / 1e3:8.3f} ms  {e.name[:80]}")
