import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """Build the C-ABI library once per session (nvcc cross-compiles without a GPU)."""
    from cdc_compression_b200 import build as B
    return B.build()


def import_variant(variant):
    """Import the drop-in ``modules`` package of one variant the way the demo scripts do (cwd-relative)."""
    sub = "epsilonparam" if variant == "eps" else "xparam"
    path = os.path.join(ROOT, "cdc_compression_b200", sub)
    for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
        del sys.modules[k]
    sys.path.insert(0, path)
    try:
        import importlib
        unet = importlib.import_module("modules.unet")
        dd = importlib.import_module("modules.denoising_diffusion")
        cm = importlib.import_module("modules.compress_modules")
    finally:
        sys.path.remove(path)
    return unet, dd, cm


def build_dropin(variant, with_context_fn=True):
    unet, dd, cm = import_variant(variant)
    if variant == "eps":
        u = unet.Unet(dim=64, channels=3, context_channels=3, dim_mults=(1, 2, 3, 4, 5, 6),
                      context_dim_mults=(1, 2, 3, 4))
        c = cm.BigCompressor(dim=64, dim_mults=(1, 2, 3, 4), hyper_dims_mults=(4, 4, 4), channels=3, out_channels=3,
                             vbr=False) if with_context_fn else None
        d = dd.GaussianDiffusion(denoise_fn=u, context_fn=c, num_timesteps=20000, loss_type="l1", clip_noise="none",
                                 vbr=False, lagrangian=0.9, pred_mode="noise", var_schedule="linear",
                                 aux_loss_weight=0, aux_loss_type="lpips")
    else:
        u = unet.Unet(dim=64, channels=3, context_channels=64, dim_mults=[1, 2, 3, 4, 5, 6],
                      context_dim_mults=[1, 2, 3, 4], embd_type="01")
        c = cm.ResnetCompressor(dim=64, dim_mults=[1, 2, 3, 4], reverse_dim_mults=[4, 3, 2, 1],
                                hyper_dims_mults=[4, 4, 4], channels=3, out_channels=64) if with_context_fn else None
        d = dd.GaussianDiffusion(denoise_fn=u, context_fn=c, ae_fn=None, num_timesteps=8193, loss_type="l2",
                                 lagrangian=0.0032, pred_mode="x", aux_loss_weight=0, aux_loss_type="lpips",
                                 var_schedule="cosine", use_loss_weight=True, loss_weight_min=5,
                                 use_aux_loss_weight_schedule=False)
    d.eval()
    return d
