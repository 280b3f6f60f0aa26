"""Python owner of a ``cdc_engine`` (include/cdc_b200.h).

PyTorch is plumbing here: it owns every tensor that crosses the C ABI (inputs, outputs, the
workspace from ``torch.empty``) and supplies the CUDA stream; all arithmetic of the denoiser
runs in libcdc_b200.so.  Nothing in this module computes on the CPU — missing library or
missing GPU raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _native
from ._native import CLIP, PRED, VARIANT, CdcConfig, CdcStepCoef


class EngineError(RuntimeError):
    pass


def native_available() -> bool:
    try:
        _native.load()
        return True
    except (ImportError, OSError, AttributeError):
        return False


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


class _nvtx:
    """NVTX range around one engine call (the build owns the tracing hooks): `ncu --nvtx --nvtx-include
    "cdc.sample_loop/"` or a timeline tool can then pick out the phases of a decode.  Costs ~100 ns per call."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


class DenoiserEngine:
    """One engine per (device, Unet instance).  ``device=None`` builds a planning-only engine
    (workspace sizes, launch counts, FLOPs, weight validation) that cannot compute."""

    def __init__(self, variant: str, dim: int, dim_mults: Sequence[int], context_dim_mults: Sequence[int],
                 channels: int, context_channels: int, device: Optional[torch.device]):
        self._lib = _native.load()
        self.variant = variant
        self.channels = channels
        self.n_context = len(context_dim_mults)
        # channel count of each context map: [context_channels, dim*m_0, dim*m_1, ...] (reference unet.py:28-30)
        self.context_widths = ([int(context_channels)] + [int(dim) * int(m) for m in context_dim_mults])[:self.n_context]
        cfg = CdcConfig()
        cfg.abi_version = _native.CDC_ABI_VERSION
        cfg.variant = VARIANT[variant]
        cfg.dim = dim
        cfg.channels = channels
        cfg.context_channels = context_channels
        if len(dim_mults) > _native.CDC_MAX_LEVELS or len(context_dim_mults) > _native.CDC_MAX_LEVELS:
            raise EngineError("too many levels")
        cfg.n_levels = len(dim_mults)
        for i, m in enumerate(dim_mults):
            cfg.dim_mults[i] = int(m)
        cfg.n_context = len(context_dim_mults)
        for i, m in enumerate(context_dim_mults):
            cfg.context_dim_mults[i] = int(m)
        if device is None:
            dev_index = -1
            self.device = None
        else:
            device = torch.device(device)
            if device.type != "cuda":
                raise EngineError("the CDC denoiser engine runs on CUDA (sm_100a) only; there is no CPU path")
            dev_index = device.index if device.index is not None else torch.cuda.current_device()
            self.device = torch.device("cuda", dev_index)
        handle = C.c_void_p()
        rc = self._lib.cdc_engine_create(C.byref(cfg), dev_index, C.byref(handle))
        if rc != 0:
            raise EngineError(f"cdc_engine_create: {self._lib.cdc_last_error(None).decode()} (rc={rc})")
        self._h = handle
        self._ws: Optional[torch.Tensor] = None
        self._ws_shape = None
        self._sched_S = 0
        self._keepalive = []

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc < 0:
            raise EngineError(f"{what}: {self._lib.cdc_last_error(self._h).decode()} (rc={rc})")
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cdc_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, B: int, H: int, W: int) -> torch.Tensor:
        if self._ws_shape != (B, H, W):
            need = self._check(self._lib.cdc_engine_workspace_bytes(self._h, B, H, W), "cdc_engine_workspace_bytes")
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws_shape = (B, H, W)
        return self._ws

    # ------------------------------------------------------------------ weights
    def load_weights(self, state_dict: Dict[str, torch.Tensor],
                     context_decoder: Optional[Dict[str, torch.Tensor]] = None):
        """``state_dict`` = the Unet's own keys.  fp32 host copies are handed
        to the engine, which repacks them to its fp16 kernel layouts.  ``context_decoder`` (optional) = the
        compressor's decoder entries under their GaussianDiffusion keys ('context_fn.dec.<i>....'): enables
        ``context_decode``."""
        if context_decoder:
            state_dict = {**state_dict, **context_decoder}
        for key, t in state_dict.items():
            t = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._check(self._lib.cdc_engine_set_weight(self._h, key.encode(), C.c_void_p(t.data_ptr()), shape,
                                                        t.dim()), f"cdc_engine_set_weight({key})")
        self._check(self._lib.cdc_engine_finalize(self._h), "cdc_engine_finalize")
        self._ws_shape = None

    # ------------------------------------------------------------------ introspection
    def workspace_bytes(self, B, H, W) -> int:
        return self._check(self._lib.cdc_engine_workspace_bytes(self._h, B, H, W), "cdc_engine_workspace_bytes")

    def launches_per_forward(self, B, H, W) -> int:
        return self._check(self._lib.cdc_engine_launches_per_forward(self._h, B, H, W), "launches_per_forward")

    def launches_per_step(self, B, H, W) -> int:
        return self._check(self._lib.cdc_engine_launches_per_step(self._h, B, H, W), "launches_per_step")

    def flops_per_forward(self, B, H, W) -> float:
        v = self._lib.cdc_engine_flops_per_forward(self._h, B, H, W)
        if v < 0:
            raise EngineError("flops_per_forward failed: " + self._lib.cdc_last_error(self._h).decode())
        return v

    def set_mainloop(self, kind: int):
        """0 = mma.sync (HMMA) kernels, 1 = tcgen05/TMA kernels for the stride-1 convolutions."""
        self._check(self._lib.cdc_engine_set_mainloop(self._h, int(kind)), "cdc_engine_set_mainloop")
        self._ws_shape = None

    def tc_ops(self, B, H, W) -> int:
        return self._check(self._lib.cdc_engine_tc_ops(self._h, B, H, W), "cdc_engine_tc_ops")

    def set_debug(self, no_reuse: bool):
        self._check(self._lib.cdc_engine_set_debug(self._h, int(no_reuse)), "cdc_engine_set_debug")
        self._ws_shape = None

    def debug_ops(self, B, H, W):
        n = self._check(self._lib.cdc_engine_num_ops(self._h, B, H, W), "num_ops")
        return [self._lib.cdc_engine_op_name(self._h, i).decode() for i in range(n)]

    def debug_read(self, op_index: int) -> Optional[torch.Tensor]:
        """fp16 NHWC output of plan op ``op_index`` from the last run, as fp32 [B,H,W,C] (CPU)."""
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        n = self._check(self._lib.cdc_engine_debug_read(self._h, op_index, None, 0, C.byref(c), C.byref(h),
                                                        C.byref(w)), "debug_read")
        if n == 0:
            return None
        out = torch.empty(n, dtype=torch.float32)
        self._check(self._lib.cdc_engine_debug_read(self._h, op_index, C.c_void_p(out.data_ptr()), n, C.byref(c),
                                                    C.byref(h), C.byref(w)), "debug_read")
        return out.reshape(-1, h.value, w.value, c.value)

    def profile_ops(self, iters: int = 5):
        """[(op name, avg ms, algorithmic FLOPs)] for the plan that ran last, each op timed alone."""
        cap = 4096
        ms = (C.c_float * cap)()
        fl = (C.c_double * cap)()
        n = self._check(self._lib.cdc_engine_profile_ops(self._h, iters, ms, fl, cap, self._stream()), "profile_ops")
        return [(self._lib.cdc_engine_op_name(self._h, i).decode(), ms[i], fl[i]) for i in range(n)]

    # ------------------------------------------------------------------ compute
    def _ctx_array(self, context: Sequence[torch.Tensor], B, H, W):
        if len(context) != self.n_context:
            raise EngineError(f"expected {self.n_context} context tensors, got {len(context)}")
        keep = []
        for l, c in enumerate(context):
            if c.device != self.device:
                raise EngineError("context tensor on the wrong device")
            if c.dim() != 4 or c.shape[0] != B or c.shape[2] != (H >> l) or c.shape[3] != (W >> l):
                raise EngineError(f"context[{l}] has shape {tuple(c.shape)}; expected spatial {(H >> l, W >> l)}")
            if c.shape[1] != self.context_widths[l]:
                raise EngineError(f"context[{l}] has {c.shape[1]} channels; this Unet expects "
                                  f"{self.context_widths[l]} (context_channels / dim * context_dim_mults)")
            keep.append(c.detach().to(torch.float32).contiguous())
        arr = (C.c_void_p * len(keep))(*[c.data_ptr() for c in keep])
        return arr, keep

    def _prep_x(self, x: torch.Tensor):
        if x.device != self.device:
            raise EngineError(f"input on {x.device}, engine on {self.device}: the denoiser runs on CUDA only")
        if x.dim() != 4 or x.shape[1] != self.channels:
            raise EngineError(f"expected [B,{self.channels},H,W], got {tuple(x.shape)}")
        return x.shape[0], x.shape[2], x.shape[3]

    def forward(self, x: torch.Tensor, time: torch.Tensor, context: Sequence[torch.Tensor]) -> torch.Tensor:
        """Unet.forward(x, time, context) -> [B,channels,H,W] fp32."""
        B, H, W = self._prep_x(x)
        xin = x.detach().to(torch.float32).contiguous()
        t = time.detach().to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        if t.numel() != B:
            raise EngineError(f"time must have {B} entries, got {t.numel()}")
        ws = self._workspace(B, H, W)
        arr, keep = self._ctx_array(context, B, H, W)
        out = torch.empty_like(xin)
        with _nvtx("cdc.unet_forward"):
            self._check(self._lib.cdc_unet_forward(self._h, _ptr(xin), _ptr(t), arr, len(keep), _ptr(out), B, H, W,
                                                   _ptr(ws), ws.numel(), self._stream()), "cdc_unet_forward")
        return out

    def set_context(self, context: Sequence[torch.Tensor], B, H, W):
        ws = self._workspace(B, H, W)
        arr, keep = self._ctx_array(context, B, H, W)
        with _nvtx("cdc.set_context"):
            self._check(self._lib.cdc_set_context(self._h, arr, len(keep), B, H, W, _ptr(ws), ws.numel(),
                                                  self._stream()), "cdc_set_context")

    def has_context_decoder(self) -> bool:
        return self._check(self._lib.cdc_engine_has_context_decoder(self._h), "cdc_engine_has_context_decoder") == 1

    def context_decode(self, q_latent: torch.Tensor, B, H, W):
        """``context_fn.decode(q_latent)`` on the engine: the four context maps land in the workspace in the layout
        the U-Net plan reads (replaces ``set_context(context_fn.decode(q_latent))``)."""
        if q_latent.device != self.device:
            raise EngineError("latent on the wrong device")
        q = q_latent.detach().to(torch.float32).contiguous()
        if q.dim() != 4 or q.shape[0] != B or q.shape[2] * 16 != H or q.shape[3] * 16 != W:
            raise EngineError(f"latent has shape {tuple(q.shape)}; expected [B={B}, C, {H // 16}, {W // 16}]")
        ws = self._workspace(B, H, W)
        with _nvtx("cdc.context_decode"):
            self._check(self._lib.cdc_context_decode(self._h, _ptr(q), B, H, W, _ptr(ws), ws.numel(), self._stream()),
                        "cdc_context_decode")

    def read_context(self, level: int, B, H, W) -> torch.Tensor:
        """Context map ``level`` as the engine holds it, as fp32 NCHW (tests)."""
        ws = self._workspace(B, H, W)
        out = torch.empty(B, self.context_widths[level], H >> level, W >> level, dtype=torch.float32, device=self.device)
        self._check(self._lib.cdc_engine_read_context(self._h, int(level), _ptr(out), B, H, W, _ptr(ws), ws.numel(),
                                                      self._stream()), "cdc_engine_read_context")
        return out

    def set_schedule(self, coefs: torch.Tensor):
        """``coefs``: [S, 8] fp32 CPU tensor, columns as in ``cdc_step_coef``."""
        coefs = coefs.detach().to(device="cpu", dtype=torch.float32).contiguous()
        assert coefs.dim() == 2 and coefs.shape[1] == 8
        S = coefs.shape[0]
        self._check(self._lib.cdc_set_schedule(self._h, C.cast(C.c_void_p(coefs.data_ptr()), C.POINTER(CdcStepCoef)),
                                               S, self._stream()), "cdc_set_schedule")
        self._sched_S = S

    def ddim_step(self, x: torch.Tensor, i: int, z: Optional[torch.Tensor], pred_mode: str, clip_mode: str):
        """In-place DDIM step at schedule index i on contiguous fp32 ``x``."""
        B, H, W = self._prep_x(x)
        assert x.dtype == torch.float32 and x.is_contiguous()
        ws = self._workspace(B, H, W)
        if z is not None:
            z = z.to(torch.float32).contiguous()
        with _nvtx("cdc.ddim_step"):
            self._check(self._lib.cdc_ddim_step(self._h, _ptr(x), int(i), _ptr(z), PRED[pred_mode], CLIP[clip_mode], B, H,
                                                W, _ptr(ws), ws.numel(), self._stream()), "cdc_ddim_step")
        return x

    def sample_loop_noise(self, x: torch.Tensor, i_first: int, i_last: int, z: torch.Tensor, pred_mode: str,
                          clip_mode: str):
        """In-place DDIM loop with eta != 0 over schedule indices i_first..i_last; ``z[j]`` is the standard-normal tensor of
        step i_first - j (drawn by the caller in loop order).  One CUDA-graph replay per step."""
        B, H, W = self._prep_x(x)
        assert x.dtype == torch.float32 and x.is_contiguous()
        n = i_first - i_last + 1
        if z.device != self.device or z.dtype != torch.float32 or not z.is_contiguous() or z.numel() < n * x.numel():
            raise EngineError("noise buffer must be a contiguous fp32 CUDA tensor holding one x-shaped tensor per step")
        ws = self._workspace(B, H, W)
        with _nvtx("cdc.sample_loop_noise"):
            self._check(self._lib.cdc_sample_loop_noise(self._h, _ptr(x), int(i_first), int(i_last), _ptr(z), PRED[pred_mode],
                                                        CLIP[clip_mode], B, H, W, _ptr(ws), ws.numel(), self._stream()),
                        "cdc_sample_loop_noise")
        return x

    def sample_loop(self, x: torch.Tensor, i_first: int, i_last: int, pred_mode: str, clip_mode: str):
        """In-place eta=0 DDIM loop over schedule indices i_first..i_last (CUDA-graph replay per step)."""
        B, H, W = self._prep_x(x)
        assert x.dtype == torch.float32 and x.is_contiguous()
        ws = self._workspace(B, H, W)
        with _nvtx("cdc.sample_loop"):
            self._check(self._lib.cdc_sample_loop(self._h, _ptr(x), int(i_first), int(i_last), PRED[pred_mode],
                                                  CLIP[clip_mode], B, H, W, _ptr(ws), ws.numel(), self._stream()),
                        "cdc_sample_loop")
        return x
