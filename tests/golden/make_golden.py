"""Generate golden vectors by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports the reference's modules from /root/reference (oracle/ref_loader.py), loads the
deterministic weights of oracle.cdc_oracle.seeded_unet_state_dict, runs
  * Unet.forward on seeded (x_t, time, context) at small sizes, both variants
  * a short DDIM p_sample_loop (S steps) from a seeded init, both variants
  * GaussianDiffusion.set_sample_schedule tables
and writes tests/golden/*.npz (fp32).  The GPU box has no /root/reference: the tests there
rebuild the same seeded inputs and compare against these files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import cdc_oracle as O  # noqa: E402
from oracle.ref_loader import build_reference_diffusion  # noqa: E402

CASES = [  # name, variant, B, H, W, seed, gain
    ("eps_b2_64x96", "eps", 2, 64, 96, 0, 1.0),
    ("x_b2_64x64", "x", 2, 64, 64, 0, 1.0),
    ("eps_b1_32x32_smallgain", "eps", 1, 32, 32, 1, 0.5),
    ("x_b1_96x64", "x", 1, 96, 64, 1, 1.0),
]
CTXDEC = [  # name, variant, B, H, W, seed   (context_fn.decode: SURVEY.md section 8 (f), row 1)
    ("eps_b1_32x64", "eps", 1, 32, 64, 0),
    ("x_b1_32x64", "x", 1, 32, 64, 1),
]
LOOPS = [  # name, variant, B, H, W, S, seed
    ("eps_loop_s6", "eps", 2, 64, 64, 6, 0),
    ("x_loop_s5", "x", 2, 64, 64, 5, 0),
]
# round 2: the benchmarked geometries (BASELINE.json configs 1-4: 256x256 tiles of 16x8 pixels, 512x768 = the demo
# scripts' Kodak input), long trajectories (50-step drift; S=65 = the x demo's default) and the
# x-variant's other pred_mode branches (xparam/modules/denoising_diffusion.py:157-165)
BIG_CASES = [  # name, variant, B, H, W, seed, gain
    ("eps_b1_256x256", "eps", 1, 256, 256, 3, 1.0),
    ("x_b1_256x256", "x", 1, 256, 256, 3, 1.0),
    ("eps_b1_512x768", "eps", 1, 512, 768, 4, 1.0),
]
LONG_LOOPS = [  # name, variant, B, H, W, S, seed, gain
    ("eps_loop_s50_smallgain", "eps", 1, 64, 64, 50, 2, 0.5),
    ("x_loop_s65", "x", 1, 64, 64, 65, 2, 1.0),
]
PRED_LOOPS = [  # name, pred_mode, B, H, W, S, seed   (x variant, clip_denoised=True)
    ("x_noise_s5", "noise", 1, 32, 32, 5, 0),
    ("x_v_s5", "v", 1, 32, 32, 5, 0),
]


def case_inputs(variant, B, H, W, seed):
    ctx = O.seeded_context(variant, B, H, W, seed=seed)
    g = torch.Generator().manual_seed(5 + seed)
    x = torch.randn(B, 3, H, W, generator=g)
    t = torch.linspace(0.15, 0.9, B)[:, None]
    init = torch.randn(B, 3, H, W, generator=g) * 0.8
    return x, t, ctx, init


def ctxdec_latent(sd, B, H, W, seed):
    """Seeded integer-valued latent with the channel count the decoder's first ResnetBlock expects (1/16 resolution)."""
    C = sd["context_fn.dec.0.0.block1.block.0.weight"].shape[1]
    g = torch.Generator().manual_seed(11 + seed)
    return (torch.randn(B, C, H // 16, W // 16, generator=g) * 2).round()


def main_round2():
    """Fixtures added in round 2 (the round-1 files above are left untouched)."""
    torch.set_grad_enabled(False)
    for name, variant, B, H, W, seed, gain in BIG_CASES:
        _, diff = build_reference_diffusion(variant, with_context_fn=False)
        diff.denoise_fn.load_state_dict(O.seeded_unet_state_dict(variant, seed, gain=gain))
        x, t, ctx, _ = case_inputs(variant, B, H, W, seed)
        y = diff.denoise_fn(x, t, ctx)
        np.savez_compressed(os.path.join(HERE, f"unet_{name}.npz"), out=y.numpy().astype(np.float32),
                            meta=np.array([B, H, W, seed], dtype=np.int64), gain=np.float32(gain))
        print(name, tuple(y.shape), float(y.abs().mean()))
    for name, variant, B, H, W, S, seed, gain in LONG_LOOPS:
        _, diff = build_reference_diffusion(variant, with_context_fn=False)
        diff.denoise_fn.load_state_dict(O.seeded_unet_state_dict(variant, seed, gain=gain))
        _, _, ctx, init = case_inputs(variant, B, H, W, seed)
        diff.set_sample_schedule(S, torch.device("cpu"))
        if variant == "eps":
            out = diff.p_sample_loop(init.shape, ctx, "ddim", init=init.clone(), eta=0)
        else:
            out = diff.p_sample_loop(init.shape, ctx, clip_denoised=True, init=init.clone(), eta=0)
        np.savez_compressed(os.path.join(HERE, f"loop_{name}.npz"), out=out.numpy(),
                            meta=np.array([B, H, W, S, seed], dtype=np.int64), gain=np.float32(gain))
        print(name, tuple(out.shape), float(out.abs().max()))
    for name, pred_mode, B, H, W, S, seed in PRED_LOOPS:
        _, diff = build_reference_diffusion("x", with_context_fn=False)
        diff.pred_mode = pred_mode
        diff.denoise_fn.load_state_dict(O.seeded_unet_state_dict("x", seed))
        _, _, ctx, init = case_inputs("x", B, H, W, seed)
        diff.set_sample_schedule(S, torch.device("cpu"))
        out = diff.p_sample_loop(init.shape, ctx, clip_denoised=True, init=init.clone(), eta=0)
        np.savez_compressed(os.path.join(HERE, f"loop_{name}.npz"), out=out.numpy(),
                            meta=np.array([B, H, W, S, seed], dtype=np.int64))
        print(name, tuple(out.shape), float(out.abs().max()))


def main():
    torch.set_grad_enabled(False)
    for name, variant, B, H, W, seed, gain in CASES:
        _, diff = build_reference_diffusion(variant, with_context_fn=False)
        diff.denoise_fn.load_state_dict(O.seeded_unet_state_dict(variant, seed, gain=gain))
        x, t, ctx, _ = case_inputs(variant, B, H, W, seed)
        y = diff.denoise_fn(x, t, ctx)
        np.savez_compressed(os.path.join(HERE, f"unet_{name}.npz"), out=y.numpy(),
                            meta=np.array([B, H, W, seed], dtype=np.int64), gain=np.float32(gain))
        print(name, tuple(y.shape), float(y.abs().mean()))
    for name, variant, B, H, W, S, seed in LOOPS:
        _, diff = build_reference_diffusion(variant, with_context_fn=False)
        diff.denoise_fn.load_state_dict(O.seeded_unet_state_dict(variant, seed))
        _, _, ctx, init = case_inputs(variant, B, H, W, seed)
        diff.set_sample_schedule(S, torch.device("cpu"))
        if variant == "eps":
            out = diff.p_sample_loop(init.shape, ctx, "ddim", init=init.clone(), eta=0)
        else:
            out = diff.p_sample_loop(init.shape, ctx, clip_denoised=True, init=init.clone(), eta=0)
        tables = {k: getattr(diff, k).numpy() for k in (
            "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod_prev", "one_minus_alphas_cumprod_prev",
            "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "sigma")}
        np.savez_compressed(os.path.join(HERE, f"loop_{name}.npz"), out=out.numpy(),
                            meta=np.array([B, H, W, S, seed], dtype=np.int64), **tables)
        print(name, tuple(out.shape), float(out.abs().max()))
    # context_fn.decode (the producer of the U-Net's context list) on a seeded integer latent
    for name, variant, B, H, W, seed in CTXDEC:
        _, diff = build_reference_diffusion(variant, with_context_fn=True)
        diff.load_state_dict(O.seeded_fill(diff.state_dict(), seed=seed, denoiser_gain=0.5))
        q = ctxdec_latent(diff.state_dict(), B, H, W, seed)
        outs = diff.context_fn.decode(q) if variant == "x" else diff.context_fn.decode(q, None)
        np.savez_compressed(os.path.join(HERE, f"ctxdec_{name}.npz"), meta=np.array([B, H, W, seed], dtype=np.int64),
                            **{f"out{i}": o.numpy() for i, o in enumerate(outs)})
        print("ctxdec", name, [tuple(o.shape) for o in outs], float(outs[0].abs().mean()))
    # schedule-only fixtures at the demo step counts
    for variant, sched, T in (("eps", "linear", 20000), ("x", "cosine", 8193)):
        _, diff = build_reference_diffusion(variant, with_context_fn=False)
        for S in (1, 65, 500):
            diff.set_sample_schedule(S, torch.device("cpu"))
            np.savez_compressed(os.path.join(HERE, f"sched_{variant}_{S}.npz"),
                                alphas_cumprod=diff.alphas_cumprod.numpy(), sigma=diff.sigma.numpy(),
                                sqrt_recipm1_alphas_cumprod=diff.sqrt_recipm1_alphas_cumprod.numpy())


if __name__ == "__main__":
    if "--round2" in sys.argv:
        main_round2()
    else:
        main()
        main_round2()
