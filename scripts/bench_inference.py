"""BASELINE.json config 5: ModelLoader.get_binary_road_map sweep over the batch size on one GPU (bf16 and fp32
paths, same random-init weights): scenes/s per batch size, and how many binarised pixels differ between the two
paths (ppm).  One JSON line per dtype; CUDA-event timing, inputs resident in HBM, 3 warm-up calls."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from driving_dirty_b200.model_loader import ModelLoader
from driving_dirty_b200.synthetic import random_roadmap_model, scene_batch

dev = torch.device("cuda:0")
batches = [int(b) for b in os.environ.get("BATCHES", "1 2 4 8 16 32 64 128 256").split()]
views_all, _ = scene_batch(max(batches), seed=20200507)
views_all = views_all.to(dev)
m32 = random_roadmap_model(dtype="fp32", device=dev)
sd = {k: v.detach().clone() for k, v in m32.state_dict().items()}
loaders = {"fp32": ModelLoader(m32, device=dev), "bf16": ModelLoader(random_roadmap_model(dtype="bf16", device=dev, state_dict=sd), device=dev)}
maps = {}
for name, loader in loaders.items():
    rows = []
    for B in batches:
        x = views_all[:B]
        iters = 20 if B <= 32 else 5
        for _ in range(3):
            torch.manual_seed(1)
            out = loader.get_binary_road_map(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            torch.manual_seed(1)          # the always-on dropout (components.py:108) draws from this stream
            out = loader.get_binary_road_map(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rows.append({"batch": B, "ms": round(ms, 4), "scenes_per_s": round(B / ms * 1e3, 1)})
        maps[(name, B)] = out
    print(json.dumps({"metric": "6-view scenes/sec, ModelLoader.get_binary_road_map (inference)", "dtype": name, "n_gpus": 1,
                      "data": "synthetic", "sweep": rows}), flush=True)
flips = []
for B in batches:
    a, b = maps[("fp32", B)], maps[("bf16", B)]
    flips.append({"batch": B, "flipped_ppm_bf16_vs_fp32": round(float((a != b).float().mean()) * 1e6, 1)})
print(json.dumps({"binarised_map_differences": flips}), flush=True)
