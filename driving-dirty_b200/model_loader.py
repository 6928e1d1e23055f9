"""ModelLoader: the competition-style inference front-end named by the north star.

The reference tree does not contain it; the interface follows the public NYU DLSP20
``model_loader.py`` template.  ``get_binary_road_map`` is DEFINED by composition of reference code:
``RoadMapBCE.forward(samples)[1].round()`` (roadmap_bce_v2.py:66-81,140), and that is what the
parity tests pin.  ``get_bounding_boxes`` has no importable reference implementation:
the interface is kept, its result is parity-unpinned.
"""
from argparse import Namespace

import torch

from . import ops
from .lightning_compat import save_checkpoint  # noqa: F401
from .roadmap_model.roadmap_bce_v2 import RoadMapBCE


def get_transform_task1():
    import torchvision
    return torchvision.transforms.ToTensor()


def get_transform_task2():
    import torchvision
    return torchvision.transforms.ToTensor()


class ModelLoader:
    team_name = "driving-dirty-b200"
    team_number = 0
    round_number = 1
    team_member = []
    contact_email = ""

    def __init__(self, model_file="roadmap_bce.ckpt", device="cuda:0"):
        """``model_file``: a ``{'state_dict', 'hparams'}`` checkpoint of RoadMapBCE (hparams must
        carry ``pretrained_path`` of the AE checkpoint, as the reference's constructor needs it),
        or an already constructed RoadMapBCE."""
        if isinstance(model_file, RoadMapBCE):
            model = model_file
        else:
            ckpt = torch.load(model_file, map_location="cpu", weights_only=False)
            hp = ckpt["hparams"]
            model = RoadMapBCE(hp if isinstance(hp, Namespace) else Namespace(**hp))
            model.load_state_dict(ckpt["state_dict"])
        self.device = torch.device(device)
        self.model = model.to(self.device)
        self.model.eval()

    @torch.no_grad()
    def get_binary_road_map(self, samples):
        """samples: CUDA tensor [B,6,3,256,306] in [0,1] -> CUDA float tensor [B,800,800] of 0./1.,
        equal to ``sigmoid(logits).round()`` of the reference forward."""
        logits = self.model._logits(samples.to(self.device))
        _, binary = ops.sigmoid_binary(logits)
        return binary.float()

    @torch.no_grad()
    def get_bounding_boxes(self, samples):
        """Tuple of B tensors [N,2,4] (metres; rows x/y; cols fl, fr, bl, br).  No reference model
        for boxes is importable; one placeholder box per sample keeps the scoring harness alive."""
        B = samples.shape[0]
        box = torch.tensor([[[2.3, 2.3, -2.3, -2.3], [1.0, -1.0, 1.0, -1.0]]], device=self.device)
        return tuple(box.clone() for _ in range(B))
