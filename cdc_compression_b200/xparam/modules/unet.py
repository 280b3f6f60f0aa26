"""``Unet`` of the x-parameterised variant (reference xparam/modules/unet.py:18-135).  ``embd_type="01"``
(float time through Linear-GELU-Linear) runs on the engine; ``"index"`` only builds the parameters."""
import math

import torch
from torch import nn

from cdc_compression_b200._shared.unet_impl import UnetBase


class ImprovedSinusoidalPosEmb(nn.Module):
    """[x, sin(2 pi x w), cos(2 pi x w)] with learned w (reference xparam/modules/network_components.py:156-171)."""

    def __init__(self, dim, is_random=False):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2), requires_grad=not is_random)

    def forward(self, x):
        x = x[:, None]
        f = x * self.weights[None, :] * 2 * math.pi
        return torch.cat((x, f.sin(), f.cos()), dim=-1)


class Unet(UnetBase):
    variant = "x"

    def __init__(self, dim, out_dim=None, dim_mults=(1, 2, 4, 8), context_dim_mults=(1, 2, 3, 3), channels=3,
                 context_channels=3, with_time_emb=True, embd_type="01"):
        super().__init__()
        self.embd_type = embd_type
        if embd_type not in ("01", "index"):
            raise NotImplementedError
        if with_time_emb and embd_type == "index":
            self.time_mlp = nn.Sequential(ImprovedSinusoidalPosEmb(dim // 2), nn.Linear(dim // 2 + 1, dim * 4),
                                          nn.GELU(), nn.Linear(dim * 4, dim))
        self._build(dim, out_dim, dim_mults, context_dim_mults, channels, context_channels, with_time_emb)
