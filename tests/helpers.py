"""Shared test / smoke / bench scaffolding (not product code)."""
import contextlib
import os
import tempfile
from argparse import Namespace

import torch
import torch.nn.functional as F

from oracle import scene_oracle as so


@contextlib.contextmanager
def cpu_rng_dropout():
    """Pin the always-on dropout (components.py:108) to the CPU generator on every device.

    F.dropout on a CPU tensor draws ``empty_like(x).bernoulli_(1-p)`` from the global CPU
    generator; this patch draws exactly that mask for CUDA tensors too, so the CUDA path and the
    CPU oracle see identical masks after the same ``torch.manual_seed``."""
    orig = F.dropout

    def dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        noise = torch.empty(x.shape, dtype=torch.float32).bernoulli_(1 - p).div_(1 - p)
        return x * noise.to(device=x.device, dtype=x.dtype)

    F.dropout = dropout
    try:
        yield
    finally:
        F.dropout = orig


def make_roadmap_model(params, hidden, latent, view_h, view_w, dtype="fp32", device="cuda:0", map_size=800):
    """Product RoadMapBCE carrying ``params`` (reference state_dict keys), built the way the
    reference demands: through a fabricated AE checkpoint (roadmap_bce_v2.py:43)."""
    from driving_dirty_b200.autoencoder.autoencoder import BasicAE, default_hparams
    from driving_dirty_b200.lightning_compat import save_checkpoint
    from driving_dirty_b200.roadmap_model.roadmap_bce_v2 import RoadMapBCE

    hp = default_hparams(hidden_dim=hidden, latent_dim=latent, input_width=6 * view_w, input_height=view_h,
                         output_width=view_w, output_height=view_h, compute_dtype=dtype)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ae.ckpt")
        ae = BasicAE(hp)
        save_checkpoint(ae, path)
        del ae
        model = RoadMapBCE(Namespace(pretrained_path=path, learning_rate=1e-3, batch_size=4,
                                     output_img_freq=10 ** 9, unfreeze_epoch_no=0, link="", compute_dtype=dtype,
                                     map_size=map_size))
    res = model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model.to(device)


def build_roadmap_pair(batch, hidden, latent, view_h, view_w, dtype="fp32", device="cuda:0", seed_w=20200505,
                       seed_x=20200506, map_size=800):
    params = so.init_roadmap_params(hidden, latent, view_h, view_w, map_hw=map_size, seed=seed_w)
    views, road = so.synthetic_scene_batch(batch, view_h, view_w, map_hw=map_size, seed=seed_x)
    model = make_roadmap_model(params, hidden, latent, view_h, view_w, dtype, device, map_size)
    return model, params, views, road


def rel_max_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max |a - ref| / max |ref|  -- the tolerance form used throughout."""
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
