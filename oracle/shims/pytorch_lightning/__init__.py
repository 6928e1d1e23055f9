"""Test-infrastructure stand-in for pytorch-lightning 0.7.5 (requirements.txt:3 of the reference).

Only what the reference's hot-path modules touch at import / construction / step time:
LightningModule.{freeze, unfreeze, load_from_checkpoint, current_epoch} and a Trainer name.
Semantics follow PL 0.7.5: freeze = requires_grad False + eval(); unfreeze = requires_grad
True + train(); load_from_checkpoint = cls(Namespace(**ckpt['hparams'])) + load_state_dict.
Used ONLY by oracle/make_golden.py to import the unmodified reference from /root/reference.
"""
from argparse import Namespace

import torch
from torch import nn


class LightningModule(nn.Module):
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
        hparams = ckpt.get("hparams", {})
        if not isinstance(hparams, Namespace):
            hparams = Namespace(**hparams)
        model = cls(hparams)
        model.load_state_dict(ckpt["state_dict"])
        return model


class Trainer:
    def __init__(self, *args, **kwargs):
        pass

    @staticmethod
    def add_argparse_args(parser):
        return parser

    @classmethod
    def from_argparse_args(cls, args, **kwargs):
        return cls()

    def fit(self, model):
        raise RuntimeError("shim Trainer cannot fit")
