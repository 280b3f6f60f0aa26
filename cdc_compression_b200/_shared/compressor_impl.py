"""Hyperprior context network (host-side PyTorch, once per image).

BASELINE.json's north_star keeps "the AE/entropy path" in PyTorch; this file gives the drop-in
``Compressor`` classes the reference's module layout (so checkpoints load) and its inference
behaviour: analysis transform -> hyper-analysis -> rounding around the learned medians / predicted
means -> estimated bits per pixel -> synthesis transform producing the 4 context maps the denoiser
consumes.  Reference: epsilonparam/modules/compress_modules.py:6-184, xparam/modules/compress_modules.py:6-177.
"""
from __future__ import annotations

import torch
from torch import nn

from .layers import FlexiblePrior, dequantize, normal_box_likelihood


def _pairs(widths):
    return list(zip(widths[:-1], widths[1:]))


class HyperpriorCompressor(nn.Module):
    """Stage containers are ``nn.ModuleList``s of ``nn.ModuleList`` rows: the first entry of a row is the
    transform, the last is the resampler / activation, anything in between is the (optional) VBR slot."""

    has_vbr_slot = False

    def _init_dims(self, dim, dim_mults, reversed_widths, hyper_dims_mults, channels, out_channels):
        self.channels = channels
        self.out_channels = out_channels
        self.dims = [channels] + [dim * m for m in dim_mults]
        self.in_out = _pairs(self.dims)
        self.reversed_dims = list(reversed_widths)
        self.reversed_in_out = _pairs(self.reversed_dims)
        self.hyper_dims = [self.dims[-1]] + [dim * m for m in hyper_dims_mults]
        self.hyper_in_out = _pairs(self.hyper_dims)
        self.reversed_hyper_dims = list(reversed([self.dims[-1] * 2] + [dim * m for m in hyper_dims_mults]))
        self.reversed_hyper_in_out = _pairs(self.reversed_hyper_dims)
        self.prior = FlexiblePrior(self.hyper_dims[-1])

    def get_extra_loss(self):
        return self.prior.get_extraloss()

    def build_network(self):
        self.enc = nn.ModuleList([])
        self.dec = nn.ModuleList([])
        self.hyper_enc = nn.ModuleList([])
        self.hyper_dec = nn.ModuleList([])

    # ---- helpers ------------------------------------------------------------------------------------
    def _run_stage(self, rows, x, cond, vbr_on_last=True):
        use_vbr = self.has_vbr_slot and getattr(self, "vbr", False)
        for i, row in enumerate(rows):
            x = row[0](x)
            if use_vbr and (vbr_on_last or i != len(rows) - 1):
                x = row[1](x, cond)
            x = row[-1](x)
        return x

    def encode(self, input, cond=None):
        latent = self._run_stage(self.enc, input, cond)
        hyper_latent = self._run_stage(self.hyper_enc, latent, cond, vbr_on_last=False)
        q_hyper_latent = dequantize(hyper_latent, self.prior.medians)
        mean, scale = self._run_stage(self.hyper_dec, q_hyper_latent, cond, vbr_on_last=False).chunk(2, 1)
        scale = scale.clamp(min=0.1)
        q_latent = dequantize(latent, mean.detach())
        state4bpp = {"latent": latent, "hyper_latent": hyper_latent, "mean": mean, "scale": scale}
        return q_latent, q_hyper_latent, state4bpp

    def decode(self, input, cond=None):
        use_vbr = self.has_vbr_slot and getattr(self, "vbr", False)
        maps = []
        for row in self.dec:
            input = row[0](input)
            if use_vbr:
                input = row[1](input, cond)
            input = row[-1](input)
            maps.append(input)
        return maps[::-1]

    def bpp(self, shape, state4bpp):
        if self.training:
            raise NotImplementedError("training-time (noise-quantised) rate estimation is out of scope")
        _, _, H, W = shape
        q_hyper = dequantize(state4bpp["hyper_latent"], self.prior.medians)
        q_latent = dequantize(state4bpp["latent"], state4bpp["mean"].detach())
        hyper_rate = -self.prior.likelihood(q_hyper).log2()
        cond_rate = -normal_box_likelihood(q_latent, state4bpp["mean"], state4bpp["scale"]).log2()
        return (hyper_rate.sum(dim=(1, 2, 3)) + cond_rate.sum(dim=(1, 2, 3))) / (H * W)

    def forward(self, input, cond=None):
        q_latent, q_hyper_latent, state4bpp = self.encode(input, cond)
        return {
            "output": self.decode(q_latent, cond),
            "bpp": self.bpp(input.shape, state4bpp),
            "q_latent": q_latent,
            "q_hyper_latent": q_hyper_latent,
        }

    # ---- shared hyper-transform rows ----------------------------------------------------------------
    def _hyper_rows(self, vbr_factory):
        n = len(self.hyper_in_out)
        for i, (a, b) in enumerate(self.hyper_in_out):
            last = i >= n - 1
            row = [nn.Conv2d(a, b, 3, 1, 1) if i == 0 else nn.Conv2d(a, b, 5, 2, 2)]
            if self.has_vbr_slot:
                row.append(vbr_factory(b) if not last else nn.Identity())
            row.append(nn.Identity() if last else nn.LeakyReLU(0.2))
            self.hyper_enc.append(nn.ModuleList(row))
        n = len(self.reversed_hyper_in_out)
        for i, (a, b) in enumerate(self.reversed_hyper_in_out):
            last = i >= n - 1
            row = [nn.Conv2d(a, b, 3, 1, 1) if last else nn.ConvTranspose2d(a, b, 5, 2, 2, 1)]
            if self.has_vbr_slot:
                row.append(vbr_factory(b) if not last else nn.Identity())
            row.append(nn.Identity() if last else nn.LeakyReLU(0.2))
            self.hyper_dec.append(nn.ModuleList(row))
