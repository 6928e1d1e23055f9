"""Synthetic scenes and random-init models for benchmarks and demos: there is no
dataset or checkpoint in the build image, so throughput is measured on random views / road maps
and default-initialised weights of the reference's architecture.  Product code: no oracle import."""
import os
import tempfile
from argparse import Namespace

import torch


def scene_batch(batch: int, view_h: int = 256, view_w: int = 306, map_hw: int = 800, seed: int = 20200505):
    """views [B,6,3,H,W] fp32 in [0,1) (ToTensor range) and a bool road map [B,map,map] (CPU tensors)."""
    g = torch.Generator().manual_seed(seed)
    views = torch.rand(batch, 6, 3, view_h, view_w, generator=g)
    road = torch.rand(batch, map_hw, map_hw, generator=g) > 0.5
    return views, road


def scene_batch_bytes(batch: int, view_h: int = 256, view_w: int = 306, map_hw: int = 800, seed: int = 20200505):
    """The same synthetic scenes as raw camera bytes: views uint8 [B,6,3,H,W] (what the JPEG decoder yields before
    ToTensor, data_helper.py:109-114) and the bool road map."""
    g = torch.Generator().manual_seed(seed)
    views = torch.randint(0, 256, (batch, 6, 3, view_h, view_w), dtype=torch.uint8, generator=g)
    road = torch.rand(batch, map_hw, map_hw, generator=g) > 0.5
    return views, road


def random_roadmap_model(hidden=256, latent=128, view_h=256, view_w=306, dtype="bf16", device="cuda:0",
                         map_size=800, seed=20200505, state_dict=None):
    """RoadMapBCE with torch-default random weights (or ``state_dict``), built the way the reference
    demands: through an AE checkpoint on disk (roadmap_bce_v2.py:43)."""
    from .autoencoder.autoencoder import BasicAE, default_hparams
    from .lightning_compat import save_checkpoint
    from .roadmap_model.roadmap_bce_v2 import RoadMapBCE

    torch.manual_seed(seed)
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
    if state_dict is not None:
        res = model.load_state_dict({k: v.clone() for k, v in state_dict.items()}, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
    return model.to(device)


def box_batch(batch: int, seed: int = 20200507):
    """Synthetic object boxes: per scene 5-20 axis-aligned 4.6 m x 2 m boxes with centres U(-30, 30) m, in
    the dataset's convention [N,2,4] (metres; rows x / y; columns fl, fr, bl, br; data_helper.py:118,129)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(batch):
        n = int(torch.randint(5, 21, (1,), generator=g))
        c = torch.rand(n, 2, generator=g) * 60 - 30
        xs = torch.stack([c[:, 0] + 2.3, c[:, 0] + 2.3, c[:, 0] - 2.3, c[:, 0] - 2.3], dim=1)
        ys = torch.stack([c[:, 1] + 1.0, c[:, 1] - 1.0, c[:, 1] + 1.0, c[:, 1] - 1.0], dim=1)
        out.append(torch.stack([xs, ys], dim=1))
    return out


def random_ae_model(hidden=256, latent=128, view_h=256, view_w=306, dtype="bf16", device="cuda:0", seed=20200505):
    """BasicAE (autoencoder.py) with torch-default random weights."""
    from .autoencoder.autoencoder import BasicAE, default_hparams
    torch.manual_seed(seed)
    hp = default_hparams(hidden_dim=hidden, latent_dim=latent, input_width=6 * view_w, input_height=view_h,
                         output_width=view_w, output_height=view_h, compute_dtype=dtype)
    return BasicAE(hp).to(device)


def random_bb_model(hidden=256, latent=128, dtype="bf16", device="cuda:0", seed=20200505):
    """BBSpatialRoadMap (spatial_w_rm.py) with torch-default random weights, built through an AE checkpoint on disk like
    the reference demands (spatial_w_rm.py:44)."""
    from .autoencoder.autoencoder import BasicAE, default_hparams
    from .bounding_box_model.spatial_bb.spatial_w_rm import BBSpatialRoadMap
    from .lightning_compat import save_checkpoint
    torch.manual_seed(seed)
    hp = default_hparams(hidden_dim=hidden, latent_dim=latent, compute_dtype=dtype)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ae.ckpt")
        save_checkpoint(BasicAE(hp), path)
        model = BBSpatialRoadMap(Namespace(pretrained_path=path, learning_rate=1e-3, batch_size=4, output_img_freq=10 ** 9,
                                           unfreeze_epoch_no=0, link="", mse_loss=False, compute_dtype=dtype))
    return model.to(device)
