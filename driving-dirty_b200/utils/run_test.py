"""Evaluation driver in the shape the reference's README advertises (``README.md:35-36``: ``src/utils/run_test.py``, which
the snapshot does not contain -- SURVEY D1): load a model through ``ModelLoader``, run both competition tasks over labelled
scenes and print the two scores the NYU DLSP20 harness reported,

  * task 1 -- ``get_bounding_boxes``  scored with ``compute_ats_bounding_boxes`` (helper.py:33-72), averaged over scenes;
  * task 2 -- ``get_binary_road_map`` scored with ``compute_ts_road_map`` (helper.py:74-77), averaged over scenes.

The DLSP20 dataset is not in the build image, so scenes come from ``--scenes FILE`` (a ``torch.save``d dict with ``samples``
[N,6,3,H,W] fp32 or uint8, ``road_images`` [N,800,800] bool, optional ``boxes`` list of [n,2,4]) or are synthesised
(``--synthetic N``).  Everything on the device path runs through libdd_b200.so; there is no CPU fallback.

    python -m driving_dirty_b200.utils.run_test --model_file roadmap.ckpt --scenes val.pt --batch_size 16
    python -m driving_dirty_b200.utils.run_test --synthetic 8            # random-init model: a smoke run of the harness
"""
import argparse
import json
import time

import torch


def load_scenes(args):
    if args.scenes:
        d = torch.load(args.scenes, map_location="cpu", weights_only=False)
        boxes = d.get("boxes")
        return d["samples"], d["road_images"].bool(), boxes
    from ..synthetic import box_batch, scene_batch, scene_batch_bytes
    make = scene_batch_bytes if args.bytes else scene_batch
    views, road = make(args.synthetic, args.view_h, args.view_w, seed=args.seed)
    return views, road, box_batch(args.synthetic, seed=args.seed + 1)


def build_loader(args):
    from ..model_loader import ModelLoader
    if args.model_file:
        return ModelLoader(args.model_file, device=args.device)
    from ..synthetic import random_roadmap_model
    model = random_roadmap_model(args.hidden_dim, args.latent_dim, args.view_h, args.view_w, dtype=args.compute_dtype,
                                 device=args.device)
    return ModelLoader(model, device=args.device)


def evaluate(loader, samples, road_images, boxes, batch_size, verbose=False):
    """Returns dict(road_map_ts, bounding_box_ats, scenes, seconds): per-scene scores averaged like the harness did."""
    from .helper import compute_ats_bounding_boxes, compute_ts_road_map
    dev = loader.device
    n = samples.shape[0]
    ts_sum = torch.zeros((), device=dev)
    ats_sum = torch.zeros((), device=dev)
    n_ats = 0
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for lo in range(0, n, batch_size):
        batch = samples[lo: lo + batch_size]
        road = road_images[lo: lo + batch_size].to(dev, non_blocking=True).float()
        pred_maps = loader.get_binary_road_map(batch)                       # host batches go through pinned staging
        pred_boxes = loader.get_bounding_boxes(batch.to(dev) if not batch.is_cuda else batch)
        for i in range(batch.shape[0]):
            ts = compute_ts_road_map(pred_maps[i], road[i])
            ts_sum += ts
            if boxes is not None:
                ats = compute_ats_bounding_boxes(pred_boxes[i], boxes[lo + i].to(dev))
                ats_sum += ats
                n_ats += 1
            if verbose:
                print(f"scene {lo + i}: road map threat score {float(ts):.4f}")
    torch.cuda.synchronize(dev)
    sec = time.perf_counter() - t0
    return {"road_map_ts": float(ts_sum) / n, "bounding_box_ats": float(ats_sum) / n_ats if n_ats else None, "scenes": n,
            "seconds": sec}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--model_file", default=None, help="RoadMapBCE checkpoint {'state_dict','hparams'}; default: random init")
    ap.add_argument("--scenes", default=None, help="torch file with samples / road_images / boxes")
    ap.add_argument("--synthetic", type=int, default=8, help="number of synthetic scenes when --scenes is not given")
    ap.add_argument("--bytes", action="store_true", help="synthetic views as raw camera bytes (uint8)")
    ap.add_argument("--batch_size", type=int, default=16)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--compute_dtype", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--hidden_dim", type=int, default=256)
    ap.add_argument("--latent_dim", type=int, default=128)
    ap.add_argument("--view_h", type=int, default=256)
    ap.add_argument("--view_w", type=int, default=306)
    ap.add_argument("--seed", type=int, default=20200505)
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("run_test: needs a CUDA device (sm_100a); the scene pipeline has no CPU fallback")
    loader = build_loader(args)
    samples, road_images, boxes = load_scenes(args)
    torch.manual_seed(args.seed)          # the encoder's dropout is always on (components.py:108): pin its stream
    res = evaluate(loader, samples, road_images, boxes, args.batch_size, args.verbose)
    print(f"{loader.team_name} - Bounding Box Score: {res['bounding_box_ats'] if res['bounding_box_ats'] is not None else float('nan'):.4} "
          f"- Road Map Score: {res['road_map_ts']:.4}")
    print(json.dumps(res))
    return res


if __name__ == "__main__":
    main()
