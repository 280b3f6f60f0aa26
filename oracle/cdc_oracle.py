"""CPU oracle for the CDC conditional-diffusion decoder hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module.
The product path (``cdc_compression_b200``) never imports it and fails loudly when the
CUDA extension is missing.

This is a *functional restatement* (state_dict in, tensors out; no ``nn.Module``) of the
reference's denoiser U-Net forward and DDIM sampler.  All citations are relative to the
upstream tree (buggyyang/CDC_compression @ 00b10de):

* ``layer_norm``        epsilonparam/modules/network_components.py:56-66
* ``block``             epsilonparam/modules/network_components.py:83-91
* ``resnet_block``      epsilonparam/modules/network_components.py:94-114
* ``linear_attention``  epsilonparam/modules/network_components.py:117-139 (+PreNorm :69-77, Residual :10-16)
* ``downsample``        epsilonparam/modules/network_components.py:45-53
* ``upsample``          epsilonparam/modules/network_components.py:34-42
* ``time_embedding``    epsilonparam/modules/unet.py:40 ; xparam/modules/unet.py:40 (embd_type "01")
* ``unet_forward``      epsilonparam/modules/unet.py:95-124 ; xparam/modules/unet.py:106-135
* ``beta schedules``    epsilonparam/modules/utils.py:50-66
* ``SampleSchedule``    epsilonparam/modules/denoising_diffusion.py:81-97 ; xparam/...:89-108
* ``ddim_step_eps``     epsilonparam/modules/denoising_diffusion.py:137-152 (+ :99-103)
* ``ddim_step_x``       xparam/modules/denoising_diffusion.py:152-174 (+ :110-114, :140-150)
* ``sample_loop``       epsilonparam/modules/denoising_diffusion.py:166-192 ; xparam/...:179-205

Parity pinning: the reference ships NO tests, golden vectors or checkpoints (SURVEY.md §4,
§8c) — "parity unpinned" by the reference's own test-suite.  The oracle is instead pinned
against the *live reference code*: ``tests/golden/make_golden.py`` imports the reference
modules from /root/reference in the build container, runs them on seeded weights/inputs
and commits the outputs as fixtures; ``tests/test_oracle.py`` checks this restatement
against those fixtures (everywhere) and against the live reference (when present).

The arithmetic is done with torch CPU tensor ops (the numpy-equivalent here); ``dtype`` may be
float32 (the reference's precision) or float64 (arbiter precision for error budgets).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------
# network components
# ----------------------------------------------------------------------------------------
def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """Per-pixel LayerNorm over the channel axis, biased variance, eps inside the sqrt."""
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mean = x.mean(dim=1, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * g + b


def block(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """Conv(k=3|7, same padding)+bias -> channel LayerNorm -> ReLU.  ``p`` = '<...>.block1.'"""
    w = sd[p + "block.0.weight"]
    k = w.shape[-1]
    y = F.conv2d(x, w, sd[p + "block.0.bias"], padding=k // 2)
    y = layer_norm(y, sd[p + "block.1.g"], sd[p + "block.1.b"])
    return F.relu(y)


def _tap(taps: Optional[dict], name: str, t: Tensor) -> Tensor:
    """Record an intermediate under the CUDA engine's op name (per-op parity tests); no effect on the arithmetic."""
    if taps is not None:
        taps[name] = t
    return t


def resnet_block(sd: StateDict, p: str, x: Tensor, temb: Optional[Tensor], taps: Optional[dict] = None) -> Tensor:
    """block1 -> (+ Linear(LeakyReLU_0.2(temb)) broadcast over pixels) -> block2 -> + res_conv(x).

    ``taps`` (optional) receives '<p>block1' = block1 output incl. the timestep shift, '<p>res_conv', and '<p>block2' =
    the block's output (block2 + residual) — the tensors the engine's launches of the same names store."""
    h = block(sd, p + "block1.", x)
    if temb is not None and (p + "mlp.1.weight") in sd:
        shift = F.linear(F.leaky_relu(temb, 0.2), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])
        h = h + shift[:, :, None, None]
    _tap(taps, p + "block1", h)
    h = block(sd, p + "block2.", h)
    if (p + "res_conv.weight") in sd:
        r = _tap(taps, p + "res_conv", F.conv2d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"]))
    else:
        r = x
    return _tap(taps, p + "block2", h + r)


def linear_attention(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """Residual(PreNorm(LayerNorm, LinearAttention(heads=1, dim_head=C))).  ``p`` = '<...>.2.' / 'mid_attn.'

    q scaled by C^-1/2 (no softmax on q); k softmax over the pixel axis; ctx = k v^T (C x C);
    out = ctx^T q; 1x1 to_out + bias; + x.
    """
    b, c, h, w = x.shape
    xn = layer_norm(x, sd[p + "fn.norm.g"], sd[p + "fn.norm.b"])
    qkv = F.conv2d(xn, sd[p + "fn.fn.to_qkv.weight"])
    q, k, v = qkv.reshape(b, 3, c, h * w).unbind(dim=1)  # each [b, c, n]
    q = q * (c ** -0.5)
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bdn,ben->bde", k, v)
    out = torch.einsum("bde,bdn->ben", ctx, q).reshape(b, c, h, w)
    out = F.conv2d(out, sd[p + "fn.fn.to_out.weight"], sd[p + "fn.fn.to_out.bias"])
    return out + x


def downsample(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """Conv2d(C, C, 3, stride 2, pad 1)."""
    return F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=2, padding=1)


def upsample(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """ConvTranspose2d(C_in, C_out, 4, stride 2, pad 1); weight layout [C_in, C_out, 4, 4]."""
    return F.conv_transpose2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=2, padding=1)


def context_decode(sd: StateDict, p: str, q_latent: Tensor, cond: Optional[Tensor] = None) -> list:
    """``context_fn.decode``: the producer of the U-Net's ``context`` list.

    BigCompressor.decode (epsilonparam/modules/compress_modules.py:74-82, layers :144-156) and ResnetCompressor.decode
    (xparam/modules/compress_modules.py:68-74, layers :142-151): per stage ResnetBlock WITHOUT time embedding ->
    [VBRCondition: x * scale(cond) + shift(cond), 1x1 convs of the scalar cond, network_components.py:304-314, only when
    the compressor was built with vbr=True] -> Upsample (ConvTranspose2d 4/2/1; the last stage changes the channel
    count).  Returns the stage outputs finest first (``output[::-1]``).  ``p`` = 'context_fn.' (or '' for a bare
    compressor state_dict)."""
    n = _count(sd, p + "dec.")
    x, outs = q_latent, []
    for i in range(n):
        x = resnet_block(sd, f"{p}dec.{i}.0.", x, None)
        up = 1
        if (f"{p}dec.{i}.1.scale.weight") in sd:          # vbr=True: dec.i = [ResnetBlock, VBRCondition, Upsample]
            c = cond.reshape(-1, 1, 1, 1).to(x.dtype)
            x = x * F.conv2d(c, sd[f"{p}dec.{i}.1.scale.weight"], sd[f"{p}dec.{i}.1.scale.bias"]) + \
                F.conv2d(c, sd[f"{p}dec.{i}.1.shift.weight"], sd[f"{p}dec.{i}.1.shift.bias"])
            up = 2
        elif (f"{p}dec.{i}.2.conv.weight") in sd:         # vbr=False keeps an nn.Identity at index 1 (no parameters)
            up = 2
        x = upsample(sd, f"{p}dec.{i}.{up}.", x)
        outs.append(x)
    return outs[::-1]


def time_embedding(sd: StateDict, time: Tensor) -> Tensor:
    """Linear(1,4*dim) -> GELU(erf) -> Linear(4*dim, dim) on a [B,1] float time."""
    t = F.linear(time, sd["time_mlp.0.weight"], sd["time_mlp.0.bias"])
    t = F.gelu(t)
    return F.linear(t, sd["time_mlp.2.weight"], sd["time_mlp.2.bias"])


def _count(sd: StateDict, prefix: str) -> int:
    idx = set()
    for k in sd:
        if k.startswith(prefix):
            idx.add(int(k[len(prefix):].split(".")[0]))
    return len(idx)


def unet_forward(sd: StateDict, x: Tensor, time: Optional[Tensor], context: Sequence[Tensor],
                 taps: Optional[dict] = None) -> Tensor:
    """Denoiser forward.  ``sd`` holds the Unet's own keys (no 'denoise_fn.' prefix).

    encode: per level  cat([x, context[l]]) if l < len(context) ; RB ; RB ; attn ; push ; downsample
            then mid_block1.   decode: mid_attn ; mid_block2 ; per level cat(x, skip.pop()) ; RB ; RB ;
            attn ; upsample ; finally LayerNorm + 7x7 conv.
    """
    temb = time_embedding(sd, time) if (time is not None and "time_mlp.0.weight" in sd) else None
    n_down = _count(sd, "downs.")
    n_up = _count(sd, "ups.")
    skips: List[Tensor] = []
    for l in range(n_down):
        p = f"downs.{l}."
        if l < len(context):
            x = torch.cat([x, context[l]], dim=1)
        x = resnet_block(sd, p + "0.", x, temb, taps)
        x = resnet_block(sd, p + "1.", x, temb, taps)
        x = _tap(taps, p + "2.out", linear_attention(sd, p + "2.", x))
        skips.append(x)
        if (p + "3.conv.weight") in sd:
            x = _tap(taps, p + "3.down", downsample(sd, p + "3.", x))
    x = resnet_block(sd, "mid_block1.", x, temb, taps)
    x = _tap(taps, "mid_attn.out", linear_attention(sd, "mid_attn.", x))
    x = resnet_block(sd, "mid_block2.", x, temb, taps)
    for l in range(n_up):
        p = f"ups.{l}."
        x = torch.cat([x, skips.pop()], dim=1)
        x = resnet_block(sd, p + "0.", x, temb, taps)
        x = resnet_block(sd, p + "1.", x, temb, taps)
        x = _tap(taps, p + "2.out", linear_attention(sd, p + "2.", x))
        if (p + "3.conv.weight") in sd:
            x = _tap(taps, p + "3.up", upsample(sd, p + "3.", x))
    x = layer_norm(x, sd["final_conv.0.g"], sd["final_conv.0.b"])
    return F.conv2d(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"], padding=3)


def sub_state_dict(sd: StateDict, prefix: str, dtype: Optional[torch.dtype] = None) -> StateDict:
    """Strip ``prefix`` (e.g. 'denoise_fn.') from the keys that carry it."""
    out = {}
    for k, v in sd.items():
        if k.startswith(prefix):
            v = v.detach()
            out[k[len(prefix):]] = v.to(dtype) if (dtype is not None and v.is_floating_point()) else v
    return out


# ----------------------------------------------------------------------------------------
# schedules and the DDIM sampler
# ----------------------------------------------------------------------------------------
def linear_beta_schedule(timesteps: int) -> np.ndarray:
    scale = 1000 / timesteps
    return np.linspace(scale * 0.0001, scale * 0.02, timesteps)


def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> np.ndarray:
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


def train_alphas_cumprod(var_schedule: str, num_timesteps: int) -> Tensor:
    """float64 numpy cumprod, cast to fp32 like the reference's registered buffer."""
    betas = cosine_beta_schedule(num_timesteps) if var_schedule == "cosine" else linear_beta_schedule(num_timesteps)
    return torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)


@dataclass
class SampleSchedule:
    """The per-step fp32 tables ``set_sample_schedule`` derives (all length S)."""
    sample_steps: int
    num_timesteps: int
    index: Tensor                      # int64 training indices
    alphas_cumprod: Tensor
    alphas_cumprod_prev: Tensor
    sqrt_alphas_cumprod_prev: Tensor
    one_minus_alphas_cumprod_prev: Tensor
    sqrt_recip_alphas_cumprod: Tensor
    sqrt_recipm1_alphas_cumprod: Tensor
    sigma: Tensor


def make_sample_schedule(acp_train: Tensor, sample_steps: int, variant: str) -> SampleSchedule:
    """``variant`` in {'eps','x'}; the x variant special-cases S==1 and writes sigma differently."""
    T = acp_train.shape[0]
    if variant == "x" and sample_steps == 1:
        idx = torch.tensor([T - 1]).long()
    else:
        idx = torch.linspace(0, T - 1, sample_steps).long()
    acp = acp_train[idx]
    acp_prev = F.pad(acp[:-1], (1, 0), value=1.0)
    if variant == "eps":
        sigma = torch.sqrt((1 - acp_prev) / (1 - acp)) * torch.sqrt(1 - acp / acp_prev)
    else:
        sigma = torch.sqrt(1.0 - acp_prev) / torch.sqrt(1.0 - acp) * torch.sqrt(1.0 - acp / acp_prev)
    return SampleSchedule(
        sample_steps=sample_steps,
        num_timesteps=T,
        index=idx,
        alphas_cumprod=acp,
        alphas_cumprod_prev=acp_prev,
        sqrt_alphas_cumprod_prev=torch.sqrt(acp_prev),
        one_minus_alphas_cumprod_prev=1.0 - acp_prev,
        sqrt_recip_alphas_cumprod=torch.sqrt(1.0 / acp),
        sqrt_recipm1_alphas_cumprod=torch.sqrt(1.0 / acp - 1),
        sigma=sigma,
    )


def unet_time(sch: SampleSchedule, i: int, variant: str, batch: int) -> Tensor:
    """The [B,1] float the sampler feeds the U-Net at loop index i: i/S (eps) or index[i]/T (x)."""
    if variant == "eps":
        v = torch.tensor(float(i)) / sch.sample_steps
    else:
        v = sch.index[i].float() / sch.num_timesteps
    return v.reshape(1, 1).repeat(batch, 1)


def ddim_update_eps(sch: SampleSchedule, i: int, x: Tensor, noise: Tensor, clip: str = "none",
                    eta: float = 0.0, z: Optional[Tensor] = None) -> Tensor:
    x0 = sch.sqrt_recip_alphas_cumprod[i] * x - sch.sqrt_recipm1_alphas_cumprod[i] * noise
    if clip == "full":
        x0 = x0.clamp(-1.0, 1.0)
    elif clip == "half":
        x0 = x0.clone()
        x0[: x0.shape[0] // 2].clamp_(-1.0, 1.0)
    out = sch.sqrt_alphas_cumprod_prev[i] * x0 + torch.sqrt(
        sch.one_minus_alphas_cumprod_prev[i] - (eta * sch.sigma[i]) ** 2) * noise
    if eta != 0 and z is not None:
        out = out + eta * sch.sigma[i] * z
    return out


def ddim_update_x(sch: SampleSchedule, i: int, x: Tensor, fx: Tensor, clip: bool = True,
                  eta: float = 0.0, z: Optional[Tensor] = None, pred_mode: str = "x") -> Tensor:
    """xparam/modules/denoising_diffusion.py:152-174; pred_mode "noise" / "v" branches :157-165 (+ :110-114, :130-139)."""
    if pred_mode == "noise":
        x0 = sch.sqrt_recip_alphas_cumprod[i] * x - sch.sqrt_recipm1_alphas_cumprod[i] * fx
    elif pred_mode == "v":
        x0 = torch.sqrt(sch.alphas_cumprod[i]) * x - torch.sqrt(1.0 - sch.alphas_cumprod[i]) * fx
    else:
        x0 = fx
    x0 = x0.clamp(-1.0, 1.0) if clip else x0
    if pred_mode == "noise":
        noise = fx
    else:
        noise = (sch.sqrt_recip_alphas_cumprod[i] * x - x0) / sch.sqrt_recipm1_alphas_cumprod[i]
    out = sch.sqrt_alphas_cumprod_prev[i] * x0 + torch.sqrt(
        (sch.one_minus_alphas_cumprod_prev[i] - (eta * sch.sigma[i]) ** 2).clamp(min=0)) * noise
    if eta != 0 and z is not None:
        out = out + eta * sch.sigma[i] * z
    return out


def sample_loop(sd_unet: StateDict, sch: SampleSchedule, variant: str, context: Sequence[Tensor],
                init: Tensor, clip=None, steps: Optional[Sequence[int]] = None,
                unet=unet_forward, pred_mode: str = "x") -> Tensor:
    """Reversed loop over the S schedule entries (``steps`` restricts it, for bounded timing)."""
    x = init
    order = list(reversed(range(sch.sample_steps))) if steps is None else list(steps)
    for i in order:
        t = unet_time(sch, i, variant, x.shape[0]).to(x.dtype)
        f = unet(sd_unet, x, t, context)
        if variant == "eps":
            x = ddim_update_eps(sch, i, x, f, clip="none" if clip is None else clip)
        else:
            x = ddim_update_x(sch, i, x, f, clip=True if clip is None else clip, pred_mode=pred_mode)
    return x


def batch_psnr(a: Tensor, b: Tensor) -> Tensor:
    """PSNR on [0,1] images per batch element (xparam/modules/trainer.py:12-16 definition)."""
    mse = ((a - b) ** 2).flatten(1).mean(dim=1)
    return 20 * torch.log10(1.0 / torch.sqrt(mse))


# ----------------------------------------------------------------------------------------
# deterministic weights for parity work (no checkpoints are available offline)
# ----------------------------------------------------------------------------------------
def unet_param_shapes(variant: str = "eps", dim: int = 64, dim_mults=(1, 2, 3, 4, 5, 6),
                      context_dim_mults=(1, 2, 3, 4), channels: int = 3,
                      context_channels: Optional[int] = None) -> Dict[str, tuple]:
    """Key -> shape of the Unet state_dict, derived from the ctor rules
    at epsilonparam/modules/unet.py:18-93."""
    if context_channels is None:
        context_channels = 3 if variant == "eps" else 64
    dims = [channels] + [dim * m for m in dim_mults]
    cdims = [context_channels] + [dim * m for m in context_dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    nres = len(in_out)
    shapes: Dict[str, tuple] = {
        "time_mlp.0.weight": (dim * 4, 1), "time_mlp.0.bias": (dim * 4,),
        "time_mlp.2.weight": (dim, dim * 4), "time_mlp.2.bias": (dim,),
    }

    def rb(p, cin, cout, k1=3):
        shapes[p + "mlp.1.weight"] = (cout, dim)
        shapes[p + "mlp.1.bias"] = (cout,)
        shapes[p + "block1.block.0.weight"] = (cout, cin, k1, k1)
        shapes[p + "block1.block.0.bias"] = (cout,)
        shapes[p + "block1.block.1.g"] = (1, cout, 1, 1)
        shapes[p + "block1.block.1.b"] = (1, cout, 1, 1)
        shapes[p + "block2.block.0.weight"] = (cout, cout, 3, 3)
        shapes[p + "block2.block.0.bias"] = (cout,)
        shapes[p + "block2.block.1.g"] = (1, cout, 1, 1)
        shapes[p + "block2.block.1.b"] = (1, cout, 1, 1)
        if cin != cout:
            shapes[p + "res_conv.weight"] = (cout, cin, 1, 1)
            shapes[p + "res_conv.bias"] = (cout,)

    def attn(p, c):
        shapes[p + "fn.norm.g"] = (1, c, 1, 1)
        shapes[p + "fn.norm.b"] = (1, c, 1, 1)
        shapes[p + "fn.fn.to_qkv.weight"] = (3 * c, c, 1, 1)
        shapes[p + "fn.fn.to_out.weight"] = (c, c, 1, 1)
        shapes[p + "fn.fn.to_out.bias"] = (c,)

    for l, (cin, cout) in enumerate(in_out):
        last = l >= nres - 1
        c0 = cin + cdims[l] if (not last and l < len(cdims) - 1) else cin
        rb(f"downs.{l}.0.", c0, cout, 7 if l == 0 else 3)
        rb(f"downs.{l}.1.", cout, cout)
        attn(f"downs.{l}.2.", cout)
        if not last:
            shapes[f"downs.{l}.3.conv.weight"] = (cout, cout, 3, 3)
            shapes[f"downs.{l}.3.conv.bias"] = (cout,)
    mid = dims[-1]
    rb("mid_block1.", mid, mid)
    attn("mid_attn.", mid)
    rb("mid_block2.", mid, mid)
    for l, (cin, cout) in enumerate(reversed(in_out[1:])):
        rb(f"ups.{l}.0.", cout * 2, cin)
        rb(f"ups.{l}.1.", cin, cin)
        attn(f"ups.{l}.2.", cin)
        shapes[f"ups.{l}.3.conv.weight"] = (cin, cin, 4, 4)
        shapes[f"ups.{l}.3.conv.bias"] = (cin,)
    shapes["final_conv.0.g"] = (1, dim, 1, 1)
    shapes["final_conv.0.b"] = (1, dim, 1, 1)
    shapes["final_conv.1.weight"] = (channels, dim, 7, 7)
    shapes["final_conv.1.bias"] = (channels,)
    return shapes


def seeded_unet_state_dict(variant: str = "eps", seed: int = 0, gain: float = 1.0, **cfg) -> StateDict:
    """Deterministic fp32 weights from the CPU generator (identical wherever torch's CPU RNG is).

    Conv/linear weights ~ U(-a, a), a = gain*sqrt(3/fan_in); biases U(-0.1,0.1); LayerNorm g in
    [0.8,1.2], b in [-0.1,0.1] so the affine terms are exercised.  ``final_conv.1`` is scaled by
    ``gain`` only (SURVEY.md §8c suggests small-gain runs for the trained-model regime).
    """
    shapes = unet_param_shapes(variant, **cfg)
    sd: StateDict = {}
    for n, (k, shp) in enumerate(sorted(shapes.items())):
        gen = torch.Generator().manual_seed(1_000_003 * (seed + 1) + n)
        u = torch.rand(shp, generator=gen) * 2 - 1
        if k.endswith(".g"):
            sd[k] = 1.0 + 0.2 * u
        elif k.endswith(".b") or k.endswith(".bias"):
            sd[k] = 0.1 * u
        else:
            if k.endswith("3.conv.weight") and len(shp) == 4 and shp[2] == 4:
                fan_in = shp[0] * 4          # transposed conv: each output sees 2x2 taps of C_in
            else:
                fan_in = int(np.prod(shp[1:]))
            sd[k] = u * (gain * math.sqrt(3.0 / fan_in))
    return sd


def seeded_context(variant: str, batch: int, h: int, w: int, seed: int = 0, dim: int = 64,
                   context_dim_mults=(1, 2, 3, 4)) -> List[Tensor]:
    """Stand-in for ``context_fn(images)['output']``: 4 smooth-ish maps with the right shapes."""
    c0 = 3 if variant == "eps" else 64
    chans = [c0] + [dim * m for m in context_dim_mults[:-1]]
    out = []
    for l, c in enumerate(chans):
        gen = torch.Generator().manual_seed(77_000 + 31 * seed + l)
        out.append(torch.randn(batch, c, h >> l, w >> l, generator=gen) * 0.5)
    return out


def seeded_fill(state_dict: StateDict, seed: int = 0, denoiser_gain: float = 1.0) -> StateDict:
    """Deterministic replacement for a whole GaussianDiffusion state_dict (denoise_fn.* AND context_fn.*), usable on
    the reference and on the drop-in alike (same keys).  Entropy-model constants (prior.affine.*.weight, prior.a.*,
    prior._medians) and the train_* schedule buffers keep their constructor values."""
    out: StateDict = {}
    for n, (k, v) in enumerate(sorted(state_dict.items())):
        keep = (k.startswith("train_") or ".prior.a." in k or k.endswith("prior._medians")
                or (".prior.affine." in k and k.endswith(".weight")) or not v.is_floating_point())
        if keep:
            out[k] = v.clone()
            continue
        gen = torch.Generator().manual_seed(7_000_003 * (seed + 1) + n)
        u = torch.rand(v.shape, generator=gen) * 2 - 1
        if k.endswith(".g"):
            out[k] = 1.0 + 0.2 * u
        elif ".prior.affine." in k and k.endswith(".bias"):
            out[k] = 0.5 * u
        elif k.endswith(".b") or k.endswith(".bias") or v.dim() < 2:
            out[k] = 0.1 * u
        else:
            shp = tuple(v.shape)
            if len(shp) == 4 and ("dec." in k or "ups." in k) and k.endswith("conv.weight") and shp[2] in (4, 5) \
                    and "hyper_enc" not in k and "enc." not in k.split("dec.")[0][-4:]:
                fan_in = shp[0] * (shp[2] * shp[3]) / 4.0        # transposed conv, stride 2
            else:
                fan_in = float(np.prod(shp[1:]))
            gain = denoiser_gain if k.startswith("denoise_fn.final_conv") else 1.0
            out[k] = u * (gain * math.sqrt(3.0 / fan_in))
    return out


def kodak_crops(images: Sequence[Tensor], size: int = 256, count: int = 8) -> Tensor:
    """`count` size x size crops on a fixed grid over the given [3,H,W] images (BASELINE config 2's batch)."""
    crops = []
    for img in images:
        _, h, w = img.shape
        for y in range(0, h - size + 1, size):
            for x in range(0, w - size + 1, size):
                crops.append(img[:, y:y + size, x:x + size])
    return torch.stack(crops[:count])
