"""Encoder / Decoder / DenseBlock with the reference's constructor signatures, attribute names
and state_dict keys (src/autoencoder/components.py), running on libdd_b200.so.

nn.Conv2d / nn.Linear / nn.BatchNorm1d submodules are kept as PARAMETER CONTAINERS (so that
``state_dict()`` and checkpoints are interchangeable with the reference, OIHW / [out,in]
layouts included); their own forward() is never called for the conv stack or the wide linears.
Constructors draw from the torch RNG in the same order as the reference's, so the same seed
gives the same initial weights.
"""
import torch
from torch import nn
from torch.nn import functional as F

from .. import ops
from .._lib import IMPL_AUTO

_DTYPES = {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16}


def resolve_dtype(d):
    if isinstance(d, torch.dtype):
        return d
    try:
        return _DTYPES[str(d).lower()]
    except KeyError:
        raise ValueError(f"compute_dtype must be one of {sorted(_DTYPES)}, got {d!r}") from None


class DenseBlock(nn.Module):
    """Linear -> BatchNorm1d -> ReLU -> F.dropout(p) (components.py:96-109).  The dropout has no
    ``training=`` argument in the reference, so it is ALWAYS active; it stays a torch call here so
    both implementations consume the same Philox stream."""

    def __init__(self, in_dim, out_dim, drop_p=0.2):
        super().__init__()
        self.drop_p = drop_p
        self.fc1 = nn.Linear(in_dim, out_dim)
        self.fc_bn = nn.BatchNorm1d(out_dim)
        self.in_dim = in_dim
        self.impl = IMPL_AUTO
        self.allow_tf32 = False       # set by the owner on the bf16 (reduced-precision) path

    def forward(self, x):
        x = ops.linear(x, self.fc1.weight, self.fc1.bias, self.impl, allow_tf32=self.allow_tf32)
        x = self.fc_bn(x)
        x = F.relu(x)
        return F.dropout(x, self.drop_p)


class Encoder(nn.Module):
    """components.py:6-52.  ``forward(x)`` takes the stitched mosaic [B,3,H,6W] like the
    reference; ``forward_views(views)`` takes the six views [B,6,3,H,W] and folds the stitch into
    the first conv's loads (same result, one pass less).  ``compute_dtype`` selects the storage
    type of the conv activations (fp32: parity path; bf16: tensor-core path); ``c3_only`` is the
    reference's early-exit attribute (:31,44-45)."""

    def __init__(self, hidden_dim, latent_dim, in_channels, input_height, input_width, compute_dtype="fp32"):
        super().__init__()
        if in_channels != 3:
            raise ValueError("the B200 scene pipeline is built for 3-channel camera views")
        self.hidden_dim = hidden_dim
        self.latent_dim = latent_dim
        self.input_height = input_height
        self.input_width = input_width
        self.in_channels = in_channels

        self.c1 = nn.Conv2d(in_channels, 32, kernel_size=3, padding=1)
        self.c2 = nn.Conv2d(32, 32, kernel_size=3, padding=1)
        self.c3 = nn.Conv2d(32, 32, kernel_size=3, stride=2, padding=1)

        self.pooling_size = 4
        conv_out_dim = self._calculate_output_dim(in_channels, input_height, input_width, self.pooling_size)

        self.fc1 = DenseBlock(conv_out_dim, hidden_dim)
        self.fc2 = DenseBlock(hidden_dim, hidden_dim)
        self.fc_z_out = nn.Linear(hidden_dim, latent_dim)

        self.c3_only = False
        self.compute_dtype = resolve_dtype(compute_dtype)
        self.impl = IMPL_AUTO

    def _calculate_output_dim(self, in_channels, input_height, input_width, pooling_size):
        # The reference pushes torch.rand(1,C,H,W) through the convs here (components.py:33-38);
        # draw the same numbers to keep the RNG stream (and thus later inits) identical, and get
        # the size in closed form: c3 halves H and W (k3 s2 p1), the flat pool keeps n // 4.
        torch.rand(1, in_channels, input_height, input_width)
        h3, w3 = (input_height - 1) // 2 + 1, (input_width - 1) // 2 + 1
        return (32 * h3 * w3) // pooling_size

    def _tail(self, feats):
        # bf16 activations: fc1 streams its weight through the tensor cores (tf32), whether the pooled features arrive as
        # bf16 (training) or as the fp32 copy of the same values (inference); owners may change compute_dtype after __init__
        self.fc1.allow_tf32 = self.compute_dtype == torch.bfloat16
        x = self.fc1(feats)
        x = self.fc2(x)
        return ops.linear(x, self.fc_z_out.weight, self.fc_z_out.bias, self.impl)

    def _stack(self, inp):
        return ops.encoder_conv_stack(inp, self.c1, self.c2, self.c3, act_dtype=self.compute_dtype,
                                      c3_only=self.c3_only, impl=self.impl)

    def forward(self, x):
        feats = self._stack(x)
        return feats if self.c3_only else self._tail(feats)

    def forward_views(self, views):
        feats = self._stack(ops.as_view_batch(views, keep_bytes=True))
        return feats if self.c3_only else self._tail(feats)


class Decoder(nn.Module):
    """components.py:55-93: latent -> DenseBlock x2 -> [B,64,h,w] -> 3 ConvTranspose2d+ReLU -> 1x1
    ConvTranspose2d.  Parameters, shapes and init order match the reference."""

    def __init__(self, hidden_dim, latent_dim, in_channels, output_height, output_width, compute_dtype="fp32"):
        super().__init__()
        self.compute_dtype = resolve_dtype(compute_dtype)
        self.deconv_dim_h, self.deconv_dim_w = self._calculate_output_size(in_channels, output_height, output_width)
        self.latent_dim = latent_dim
        self.fc1 = DenseBlock(latent_dim, hidden_dim)
        self.fc2 = DenseBlock(hidden_dim, self.deconv_dim_h * self.deconv_dim_w * 64)
        self.fc2.allow_tf32 = self.compute_dtype == torch.bfloat16
        self.dc1 = nn.ConvTranspose2d(64, 32, kernel_size=3, padding=1)
        self.dc2 = nn.ConvTranspose2d(32, 32, kernel_size=3, padding=1)
        self.dc3 = nn.ConvTranspose2d(32, 32, kernel_size=2, stride=2)
        self.dc4 = nn.ConvTranspose2d(32, in_channels, kernel_size=1, stride=1)

    def _calculate_output_size(self, in_channels, output_height, output_width):
        # The reference builds four throw-away Conv2d layers and runs a random image through them
        # (components.py:75-83).  Re-draw the same RNG values (input + four default inits) and
        # compute the size directly: only the k2 s2 conv changes it.
        torch.rand(1, in_channels, output_height, output_width)
        nn.Conv2d(in_channels, 32, kernel_size=1, stride=1)
        nn.Conv2d(32, 32, kernel_size=2, stride=2)
        nn.Conv2d(32, 32, kernel_size=3, padding=1)
        nn.Conv2d(32, 64, kernel_size=3, padding=1)
        return (output_height - 2) // 2 + 1, (output_width - 2) // 2 + 1

    def forward(self, z):
        x = self.fc1(z)
        x = self.fc2(x)
        x = x.view(x.size(0), 64, self.deconv_dim_h, self.deconv_dim_w)
        return ops.decoder_deconv_stack(x, self.dc1, self.dc2, self.dc3, self.dc4, act_dtype=self.compute_dtype)
