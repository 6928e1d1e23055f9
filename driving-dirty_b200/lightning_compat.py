"""LightningModule base for the drop-in task modules.

The reference subclasses pytorch_lightning.LightningModule (0.7.5, requirements.txt:3) and uses
four things from it on this path: ``freeze`` / ``unfreeze`` (roadmap_bce_v2.py:46,129),
``load_from_checkpoint`` (:43) and ``current_epoch`` (:127).  When pytorch-lightning is installed
the real class is used; otherwise this minimal base supplies those four with PL 0.7.5 semantics,
so the modules work as plain nn.Modules driven by any loop.
"""
from argparse import Namespace

import torch
from torch import nn

try:  # pragma: no cover - not installed in the build image
    from pytorch_lightning import LightningModule as _PLModule  # type: ignore
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _PLModule = None
    HAVE_LIGHTNING = False


class _MinimalLightningModule(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.current_epoch = 0
        self.global_step = 0
        self.logger = None
        self.trainer = None

    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def unfreeze(self):
        for p in self.parameters():
            p.requires_grad = True
        self.train()

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        hparams = ckpt.get("hparams", ckpt.get("hyper_parameters", {}))
        if not isinstance(hparams, Namespace):
            hparams = Namespace(**hparams)
        model = cls(hparams)
        model.load_state_dict(ckpt["state_dict"])
        return model


LightningModule = _PLModule if HAVE_LIGHTNING else _MinimalLightningModule


def save_checkpoint(module, path):
    """Write the ``{'state_dict', 'hparams'}`` file load_from_checkpoint reads (PL 0.7.5 layout)."""
    hp = getattr(module, "hparams", None)
    # clones: a parameter adopted by optim.FusedAdam's sharded form lives inside one large symmetric buffer, and
    # torch.save writes a tensor's whole storage
    sd = {k: v.detach().clone() for k, v in module.state_dict().items()}
    torch.save({"state_dict": sd, "hparams": vars(hp) if isinstance(hp, Namespace) else dict(hp or {})}, path)
