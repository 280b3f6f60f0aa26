"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  /root/reference does not exist on the GPU box, so
nothing that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may call this; it is
used by ``tests/golden/make_golden.py`` (fixture generation) and by the CPU tests that check
the oracle restatement and the drop-in modules against the live reference (skipped when the
tree is absent).

Both variants name their package ``modules``; they are loaded here under the aliases
``cdc_ref_eps`` / ``cdc_ref_x`` so they can coexist in one process.  ``lpips`` (imported at
module top by denoising_diffusion.py:7 but only used when aux_loss_weight>0) is stubbed.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CDC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "epsilonparam", "modules"))


def _load_package(alias: str, path: str):
    if alias in sys.modules:
        return sys.modules[alias]
    sys.modules.setdefault("lpips", types.ModuleType("lpips"))
    spec = importlib.util.spec_from_file_location(
        alias, os.path.join(path, "__init__.py"), submodule_search_locations=[path])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules[alias] = pkg
    spec.loader.exec_module(pkg)
    return pkg


def load_reference(variant: str):
    """Returns a namespace with Unet, GaussianDiffusion, compressors and network_components."""
    assert variant in ("eps", "x")
    sub = "epsilonparam" if variant == "eps" else "xparam"
    alias = "cdc_ref_eps" if variant == "eps" else "cdc_ref_x"
    _load_package(alias, os.path.join(REFERENCE_ROOT, sub, "modules"))
    ns = types.SimpleNamespace()
    ns.unet = importlib.import_module(alias + ".unet")
    ns.nc = importlib.import_module(alias + ".network_components")
    ns.dd = importlib.import_module(alias + ".denoising_diffusion")
    ns.cm = importlib.import_module(alias + ".compress_modules")
    ns.utils = importlib.import_module(alias + ".utils")
    ns.Unet = ns.unet.Unet
    ns.GaussianDiffusion = ns.dd.GaussianDiffusion
    return ns


def build_reference_diffusion(variant: str, with_context_fn: bool = True):
    """The exact demo configurations (epsilonparam/test_epsilonparam.py:27-56,
    xparam/test_xparam.py:29-61) with aux_loss_weight=0 (no LPIPS)."""
    ref = load_reference(variant)
    if variant == "eps":
        unet = ref.Unet(dim=64, channels=3, context_channels=3, dim_mults=(1, 2, 3, 4, 5, 6),
                        context_dim_mults=(1, 2, 3, 4))
        ctx = ref.cm.BigCompressor(dim=64, dim_mults=(1, 2, 3, 4), hyper_dims_mults=(4, 4, 4),
                                   channels=3, out_channels=3, vbr=False) if with_context_fn else None
        diff = ref.GaussianDiffusion(denoise_fn=unet, context_fn=ctx, num_timesteps=20000, loss_type="l1",
                                     clip_noise="none", vbr=False, lagrangian=0.9, pred_mode="noise",
                                     var_schedule="linear", aux_loss_weight=0, aux_loss_type="lpips")
    else:
        unet = ref.Unet(dim=64, channels=3, context_channels=64, dim_mults=[1, 2, 3, 4, 5, 6],
                        context_dim_mults=[1, 2, 3, 4], embd_type="01")
        ctx = ref.cm.ResnetCompressor(dim=64, dim_mults=[1, 2, 3, 4], reverse_dim_mults=[4, 3, 2, 1],
                                      hyper_dims_mults=[4, 4, 4], channels=3,
                                      out_channels=64) if with_context_fn else None
        diff = ref.GaussianDiffusion(denoise_fn=unet, context_fn=ctx, ae_fn=None, num_timesteps=8193,
                                     loss_type="l2", lagrangian=0.0032, pred_mode="x", aux_loss_weight=0,
                                     aux_loss_type="lpips", var_schedule="cosine", use_loss_weight=True,
                                     loss_weight_min=5, use_aux_loss_weight_schedule=False)
    diff.eval()
    return ref, diff
