"""Parameter containers + PyTorch forwards for the blocks the *context network* uses.

The denoiser U-Net never calls these forwards: its blocks run inside libcdc_b200.so and the
modules below only own the ``nn.Parameter``s so that ``state_dict()`` keeps the reference's
key names.  The context network (``BigCompressor`` /
``ResnetCompressor``; reference ``*/modules/compress_modules.py``) runs once per image and,
as BASELINE.json's north_star states, stays host-side PyTorch — it reuses ``ResnetBlock``,
``Upsample`` and ``Downsample`` through the forwards defined here.

Behavioural reference: epsilonparam/modules/network_components.py:10-16 (Residual), :34-53
(Upsample/Downsample), :56-66 (LayerNorm), :69-77 (PreNorm), :83-114 (Block/ResnetBlock),
:117-139 (LinearAttention), :304-314 (VBRCondition), :317-412 (GDN/GDN1), :415-549 (prior).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class EngineOnly(RuntimeError):
    def __init__(self, what):
        super().__init__(f"{what} has no PyTorch forward in cdc_compression_b200: it executes inside the CUDA "
                         "engine via Unet.forward (no CPU / eager fallback)")


class LayerNorm(nn.Module):
    """Channel LayerNorm per pixel: biased variance, eps inside the sqrt, affine g/b of shape [1,C,1,1]."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))
        self.b = nn.Parameter(torch.zeros(1, dim, 1, 1))

    def forward(self, x):
        mu = x.mean(dim=1, keepdim=True)
        var = x.var(dim=1, unbiased=False, keepdim=True)
        return (x - mu) / torch.sqrt(var + self.eps) * self.g + self.b


class Block(nn.Module):
    def __init__(self, dim, dim_out, large_filter=False):
        super().__init__()
        k = 7 if large_filter else 3
        self.block = nn.Sequential(nn.Conv2d(dim, dim_out, k, padding=k // 2), LayerNorm(dim_out), nn.ReLU())

    def forward(self, x):
        return self.block(x)


class ResnetBlock(nn.Module):
    """block1 -> (+ Linear(LeakyReLU(temb))) -> block2 -> + res_conv(x)."""

    def __init__(self, dim, dim_out, time_emb_dim=None, large_filter=False):
        super().__init__()
        self.mlp = None if time_emb_dim is None else nn.Sequential(nn.LeakyReLU(0.2),
                                                                   nn.Linear(time_emb_dim, dim_out))
        self.block1 = Block(dim, dim_out, large_filter)
        self.block2 = Block(dim_out, dim_out)
        self.res_conv = nn.Identity() if dim == dim_out else nn.Conv2d(dim, dim_out, 1)

    def forward(self, x, time_emb=None):
        h = self.block1(x)
        if time_emb is not None:
            h = h + self.mlp(time_emb)[:, :, None, None]
        return self.block2(h) + self.res_conv(x)


class Upsample(nn.Module):
    def __init__(self, dim_in, dim_out=None):
        super().__init__()
        self.conv = nn.ConvTranspose2d(dim_in, dim_in if dim_out is None else dim_out, 4, 2, 1)

    def forward(self, x):
        return self.conv(x)


class Downsample(nn.Module):
    def __init__(self, dim_in, dim_out=None):
        super().__init__()
        self.conv = nn.Conv2d(dim_in, dim_in if dim_out is None else dim_out, 3, 2, 1)

    def forward(self, x):
        return self.conv(x)


class LinearAttention(nn.Module):
    """heads=1 linear attention; parameters only (to_qkv has no bias)."""

    def __init__(self, dim, heads=1, dim_head=None):
        super().__init__()
        dim_head = dim if dim_head is None else dim_head
        if heads != 1 or dim_head != dim:
            raise NotImplementedError("the CUDA engine implements heads=1, dim_head=dim (the reference's only use)")
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Conv2d(dim, dim_head * heads * 3, 1, bias=False)
        self.to_out = nn.Conv2d(dim_head * heads, dim, 1)

    def forward(self, x):
        raise EngineOnly("LinearAttention")


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = LayerNorm(dim)

    def forward(self, x):
        raise EngineOnly("PreNorm")


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, *args, **kwargs):
        raise EngineOnly("Residual")


# ------------------------------------------------------------------------------------------------
# entropy-model pieces of the context network (inference only; the training-time straight-through
# estimators of the reference are out of scope)
# ------------------------------------------------------------------------------------------------
class VBRCondition(nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.scale = nn.Conv2d(input_dim, output_dim, 1)
        self.shift = nn.Conv2d(input_dim, output_dim, 1)

    def forward(self, x, cond):
        cond = cond.reshape(-1, 1, 1, 1)
        return x * self.scale(cond) + self.shift(cond)


class GDN1(nn.Module):
    """y = x / (beta + sum_j gamma_ij |x_j|)   (or x * (...) when inverse); reparametrised like the reference."""

    def __init__(self, ch, inverse=False, beta_min=1e-6, gamma_init=0.1, reparam_offset=2 ** -18):
        super().__init__()
        self.inverse = inverse
        self.pedestal = reparam_offset ** 2
        self.beta_bound = (beta_min + self.pedestal) ** 0.5
        self.gamma_bound = reparam_offset
        self.beta = nn.Parameter(torch.sqrt(torch.ones(ch) + self.pedestal))
        self.gamma = nn.Parameter(torch.sqrt(gamma_init * torch.eye(ch) + self.pedestal))

    def forward(self, x):
        ch = x.shape[1]
        beta = self.beta.clamp(min=self.beta_bound) ** 2 - self.pedestal
        gamma = (self.gamma.clamp(min=self.gamma_bound) ** 2 - self.pedestal).view(ch, ch, 1, 1)
        norm = F.conv2d(x.abs(), gamma, beta)
        return x * norm if self.inverse else x / norm


class PriorFunction(nn.Module):
    def __init__(self, parallel_dims, in_features, out_features, scale):
        super().__init__()
        self.weight = nn.Parameter(torch.full((parallel_dims, 1, 1, in_features, out_features), float(scale)))
        self.bias = nn.Parameter(torch.empty(parallel_dims, 1, 1, 1, out_features).uniform_(-0.5, 0.5))

    def forward(self, x):
        return torch.matmul(x, F.softplus(self.weight)) + self.bias


class FlexiblePrior(nn.Module):
    """Factorised density of Ballé et al. 2018 (App. 6.1); ``likelihood`` = mass of the unit box around x."""

    def __init__(self, channels=256, dims=(3, 3, 3), init_scale=10.0):
        super().__init__()
        widths = [1, *dims, 1]
        n = len(widths) - 1
        scale = init_scale ** (1 / n)
        self.chain_len = n
        self.affine = nn.ModuleList(
            PriorFunction(channels, widths[i], widths[i + 1], np.log(np.expm1(1 / scale / widths[i + 1])))
            for i in range(n))
        self.a = nn.ParameterList(nn.Parameter(torch.zeros(channels, 1, 1, 1, widths[i + 1])) for i in range(n - 1))
        self._medians = nn.Parameter(torch.zeros(1, channels, 1, 1))

    @property
    def medians(self):
        return self._medians.detach()

    def _logits(self, x):
        y = x.transpose(0, 1).unsqueeze(-1)          # [C, B, H, W, 1]
        for i in range(self.chain_len - 1):
            y = self.affine[i](y)
            y = y + torch.tanh(self.a[i]) * torch.tanh(y)
        return self.affine[-1](y).squeeze(-1).transpose(0, 1)

    def get_extraloss(self):
        return self._logits(self._medians.detach()).detach().abs().sum()

    def likelihood(self, x, min=1e-9):
        lo, hi = self._logits(x - 0.5), self._logits(x + 0.5)
        sign = -torch.sign(lo + hi).detach()
        return (torch.sigmoid(hi * sign) - torch.sigmoid(lo * sign)).abs().clamp(min=min)


def dequantize(x, offset):
    """round(x - offset) + offset   (reference utils.quantize(mode='dequantize'))."""
    return torch.round(x - offset) + offset


def normal_box_likelihood(x, loc, scale, min=1e-9):
    """P(|X - loc| within the unit box around x) for X ~ N(loc, scale)  (reference utils.py:156-160)."""
    d = (x - loc).abs()
    cdf = lambda v: 0.5 * torch.erfc(-(2 ** -0.5) * v)
    return (cdf((0.5 - d) / scale) - cdf((-0.5 - d) / scale)).clamp(min=min)
