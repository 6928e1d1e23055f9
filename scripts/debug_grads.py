import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import scene_oracle as so
from tests.helpers import build_roadmap_pair, cpu_rng_dropout, rel_max_err
for dtype in ("fp32", "bf16"):
    for (B, hid, lat, vh, vw) in ((3, 16, 8, 16, 20), (2, 24, 8, 10, 14)):
        model, params, views, road = build_roadmap_pair(B, hid, lat, vh, vw, dtype=dtype)
        batch = (tuple(views.cuda().unbind(0)), None, tuple(road.cuda().unbind(0)))
        with cpu_rng_dropout():
            torch.manual_seed(1234)
            out = model.training_step(batch, 1)
            out["loss"].backward()
            ref, grads = so.train_step_grads(params, views, road, seed=1234)
        print(dtype, (B, vh, vw), "loss", float(out["loss"]), float(ref["loss"]))
        for k, p in model.named_parameters():
            g, r = p.grad.cpu().double(), grads[k].double()
            print(f"   {k:40s} relmax {rel_max_err(p.grad, grads[k]):.3e} relfro {float((g-r).norm()/r.norm()):.3e} norm {float(r.norm()):.3e}")
