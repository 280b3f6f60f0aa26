"""Drop-in ``Unet`` whose forward runs on the CUDA engine.

Same constructor arguments, module tree and ``state_dict`` keys as the reference
(epsilonparam/modules/unet.py:18-93, xparam/modules/unet.py:19-104); ``forward(x, time, context)``
(:120-124 / :131-135) hands the tensors to libcdc_b200.so.  The module tree exists to own the
``nn.Parameter``s (checkpoint ABI) — none of the U-Net's PyTorch modules is ever called.
"""
from __future__ import annotations

import torch
from torch import nn

from ..engine import DenoiserEngine, EngineError
from .layers import Downsample, LayerNorm, LinearAttention, PreNorm, Residual, ResnetBlock, Upsample


def _mark_engine_dirty(module, incompatible_keys):
    module._engine_dirty = True


class UnetBase(nn.Module):
    variant = "eps"

    def _build(self, dim, out_dim, dim_mults, context_dim_mults, channels, context_channels, with_time_emb):
        self.channels = channels
        self._cfg = dict(dim=dim, dim_mults=tuple(int(m) for m in dim_mults),
                         context_dim_mults=tuple(int(m) for m in context_dim_mults), channels=channels,
                         context_channels=context_channels)
        widths = [channels] + [dim * m for m in dim_mults]
        ctx_widths = [context_channels] + [dim * m for m in context_dim_mults]
        pairs = list(zip(widths[:-1], widths[1:]))
        levels = len(pairs)
        time_dim = dim if with_time_emb else None
        if not with_time_emb:
            self.time_mlp = None
        elif getattr(self, "time_mlp", None) is None:
            self.time_mlp = nn.Sequential(nn.Linear(1, dim * 4), nn.GELU(), nn.Linear(dim * 4, dim))

        def attn(c):
            return Residual(PreNorm(c, LinearAttention(c)))

        self.downs = nn.ModuleList()
        self.ups = nn.ModuleList()
        for lvl, (c_in, c_out) in enumerate(pairs):
            last = lvl >= levels - 1
            takes_ctx = (not last) and lvl < len(ctx_widths) - 1
            self.downs.append(nn.ModuleList([
                ResnetBlock(c_in + ctx_widths[lvl] if takes_ctx else c_in, c_out, time_dim, lvl == 0),
                ResnetBlock(c_out, c_out, time_dim),
                attn(c_out),
                nn.Identity() if last else Downsample(c_out),
            ]))
        mid = widths[-1]
        self.mid_block1 = ResnetBlock(mid, mid, time_dim)
        self.mid_attn = attn(mid)
        self.mid_block2 = ResnetBlock(mid, mid, time_dim)
        for c_in, c_out in reversed(pairs[1:]):
            self.ups.append(nn.ModuleList([
                ResnetBlock(c_out * 2, c_in, time_dim),
                ResnetBlock(c_in, c_in, time_dim),
                attn(c_in),
                Upsample(c_in),
            ]))
        self._out_dim = channels if out_dim is None else out_dim
        self.final_conv = nn.Sequential(LayerNorm(dim), nn.Conv2d(dim, self._out_dim, 7, padding=3))
        self._with_time_emb = with_time_emb
        self._engine = None
        self._engine_dirty = True
        # the hook receives the module it fires on, so a deep copy (ema_pytorch.EMA) marks ITSELF dirty, not the original
        self.register_load_state_dict_post_hook(_mark_engine_dirty)

    def __getstate__(self):
        """copy.deepcopy / pickle: the copy owns no engine (a ``cdc_engine`` handle must not be shared or destroyed
        twice) and re-uploads its own weights on first use."""
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_engine_dirty"] = True
        state.pop("_ctxdec_key", None)      # closes over the ORIGINAL's context_fn
        state.pop("_ctxdec_fn", None)
        return state

    # ---- engine lifecycle -------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._engine_dirty = True          # .to()/.cuda()/.half() move or change the parameters
        return super()._apply(fn, *a, **k)

    def refresh_engine(self):
        """Call after editing parameters in place (load_state_dict / .to() are tracked automatically)."""
        self._engine_dirty = True

    def engine_for(self, device, context_decoder=None) -> DenoiserEngine:
        """``context_decoder`` = (version key, callable returning the 'context_fn.dec.*' state entries) from the owning
        GaussianDiffusion: the engine then also runs ``context_fn.decode``.  Sticky: callers that pass None keep whatever
        decoder the engine already holds."""
        device = torch.device(device)
        if device.type != "cuda":
            raise EngineError("Unet.forward runs on the CUDA engine only (sm_100a); move the module and its "
                              "inputs to a GPU — there is no CPU or eager PyTorch path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if not self._with_time_emb:
            raise NotImplementedError("with_time_emb=False is not implemented by the CUDA engine")
        if self._out_dim != self.channels:
            raise NotImplementedError("out_dim != channels is not implemented by the CUDA engine")
        if getattr(self, "embd_type", "01") != "01":
            raise NotImplementedError("embd_type='index' is not implemented by the CUDA engine (SURVEY.md §2 #2)")
        if self._engine is None or self._engine.device != device:
            if self._engine is not None:
                self._engine.close()
            c = self._cfg
            self._engine = DenoiserEngine(self.variant, c["dim"], c["dim_mults"], c["context_dim_mults"],
                                          c["channels"], c["context_channels"], device)
            self._engine_dirty = True
        if context_decoder is not None and context_decoder[0] != getattr(self, "_ctxdec_key", None):
            self._engine_dirty = True
        if self._engine_dirty:
            if context_decoder is not None:
                self._ctxdec_key, self._ctxdec_fn = context_decoder
            fn = getattr(self, "_ctxdec_fn", None)
            self._engine.load_weights(self.state_dict(), fn() if fn is not None else None)
            self._engine_dirty = False
        return self._engine

    # ---- the reference interface ---------------------------------------------------------------
    def encode(self, x, t, context):
        raise NotImplementedError("encode/decode are fused inside the CUDA engine; call forward()")

    def decode(self, x, h, t):
        raise NotImplementedError("encode/decode are fused inside the CUDA engine; call forward()")

    @torch.no_grad()
    def forward(self, x, time=None, context=None):
        if time is None:
            raise NotImplementedError("time=None (no timestep embedding) is not implemented by the CUDA engine")
        eng = self.engine_for(x.device)
        return eng.forward(x, time, list(context) if context is not None else [])
