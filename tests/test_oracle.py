"""The CPU oracle (oracle/cdc_oracle.py) against the committed golden vectors — which were produced by
the unmodified reference (tests/golden/make_golden.py) — and against the live reference when present."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import cdc_oracle as O
from oracle.ref_loader import reference_available

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
from golden.make_golden import CASES, LOOPS, case_inputs  # noqa: E402

torch.set_grad_enabled(False)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_unet_forward_matches_reference_golden(case):
    name, variant, B, H, W, seed, gain = case
    gold = np.load(os.path.join(GOLD, f"unet_{name}.npz"))
    sd = O.seeded_unet_state_dict(variant, seed, gain=gain)
    x, t, ctx, _ = case_inputs(variant, B, H, W, seed)
    y = O.unet_forward(sd, x, t, ctx)
    ref = torch.from_numpy(gold["out"])
    # same fp32 ops as the reference; allow for thread-count-dependent summation order in conv kernels
    assert (y - ref).norm() / ref.norm() < 2e-6


@pytest.mark.parametrize("case", LOOPS, ids=[c[0] for c in LOOPS])
def test_sample_loop_matches_reference_golden(case):
    name, variant, B, H, W, S, seed = case
    gold = np.load(os.path.join(GOLD, f"loop_{name}.npz"))
    sd = O.seeded_unet_state_dict(variant, seed)
    _, _, ctx, init = case_inputs(variant, B, H, W, seed)
    T, sched = (20000, "linear") if variant == "eps" else (8193, "cosine")
    sch = O.make_sample_schedule(O.train_alphas_cumprod(sched, T), S, variant)
    for k in ("alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod_prev", "one_minus_alphas_cumprod_prev",
              "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "sigma"):
        a, b = getattr(sch, k).numpy(), gold[k]
        assert np.array_equal(np.nan_to_num(a), np.nan_to_num(b)), k   # bit-exact tables
    out = O.sample_loop(sd, sch, variant, ctx, init.clone())
    ref = torch.from_numpy(gold["out"])
    assert (out - ref).norm() / ref.norm() < 1e-4   # eps random-weight trajectories amplify by ~1e2


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "sched_*.npz"))), ids=os.path.basename)
def test_schedules_bit_exact(path):
    _, variant, S = os.path.basename(path)[:-4].split("_")
    gold = np.load(path)
    T, sched = (20000, "linear") if variant == "eps" else (8193, "cosine")
    sch = O.make_sample_schedule(O.train_alphas_cumprod(sched, T), int(S), variant)
    assert np.array_equal(sch.alphas_cumprod.numpy(), gold["alphas_cumprod"])
    assert np.array_equal(np.nan_to_num(sch.sigma.numpy()), np.nan_to_num(gold["sigma"]))
    assert np.array_equal(sch.sqrt_recipm1_alphas_cumprod.numpy(), gold["sqrt_recipm1_alphas_cumprod"])


def test_schedule_edge_values():
    # SURVEY.md Appendix C probes
    eps = O.make_sample_schedule(O.train_alphas_cumprod("linear", 20000), 500, "eps")
    assert abs(eps.alphas_cumprod[-1].item() - 4.30e-5) < 2e-7
    x = O.make_sample_schedule(O.train_alphas_cumprod("cosine", 8193), 65, "x")
    assert x.alphas_cumprod_prev[0].item() == 1.0
    assert torch.isfinite(x.sqrt_recipm1_alphas_cumprod).all() and x.sqrt_recipm1_alphas_cumprod[0] > 0


def test_param_shapes_counts():
    # 39 912 707 (eps) / 40 107 907 (x) U-Net parameters
    for variant, n in (("eps", 39_912_707), ("x", 40_107_907)):
        shapes = O.unet_param_shapes(variant)
        assert sum(int(np.prod(s)) for s in shapes.values()) == n


def test_ddim_update_algebra():
    sch = O.make_sample_schedule(O.train_alphas_cumprod("cosine", 8193), 9, "x")
    g = torch.Generator().manual_seed(0)
    x, f = torch.randn(2, 3, 8, 8, generator=g), torch.randn(2, 3, 8, 8, generator=g)
    # at i == 0 acp_prev == 1, so the x-variant returns the clamped prediction exactly
    assert torch.equal(O.ddim_update_x(sch, 0, x, f), f.clamp(-1, 1))
    # eps: "half" clips only the first half of the batch
    sch_e = O.make_sample_schedule(O.train_alphas_cumprod("linear", 20000), 9, "eps")
    full = O.ddim_update_eps(sch_e, 5, x, f, clip="full")
    none = O.ddim_update_eps(sch_e, 5, x, f, clip="none")
    half = O.ddim_update_eps(sch_e, 5, x, f, clip="half")
    assert torch.equal(half[:1], full[:1]) and torch.equal(half[1:], none[1:])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("variant", ["eps", "x"])
def test_oracle_equals_live_reference(variant):
    from oracle.ref_loader import build_reference_diffusion
    _, diff = build_reference_diffusion(variant, with_context_fn=False)
    sd = O.seeded_unet_state_dict(variant, 3)
    diff.denoise_fn.load_state_dict(sd)
    x, t, ctx, init = case_inputs(variant, 1, 32, 64, 3)
    assert torch.equal(diff.denoise_fn(x, t, ctx), O.unet_forward(sd, x, t, ctx))
    # every hot block in isolation, fp64
    ref, _ = build_reference_diffusion(variant, with_context_fn=False)
    nc = ref.nc
    g = torch.Generator().manual_seed(1)
    blk = nc.ResnetBlock(128, 192, 64).double()
    sdb = {"p." + k: v for k, v in blk.state_dict().items()}
    xx, te = torch.randn(2, 128, 8, 8, generator=g).double(), torch.randn(2, 64, generator=g).double()
    assert torch.allclose(blk(xx, te), O.resnet_block(sdb, "p.", xx, te), atol=1e-12)
    att = nc.Residual(nc.PreNorm(64, nc.LinearAttention(64))).double()
    sda = {"p." + k: v for k, v in att.state_dict().items()}
    xa = torch.randn(2, 64, 8, 4, generator=g).double()
    assert torch.allclose(att(xa), O.linear_attention(sda, "p.", xa), atol=1e-12)
    up, dn = nc.Upsample(64).double(), nc.Downsample(64).double()
    assert torch.allclose(up(xa), O.upsample({"p." + k: v for k, v in up.state_dict().items()}, "p.", xa), atol=1e-12)
    assert torch.allclose(dn(xa), O.downsample({"p." + k: v for k, v in dn.state_dict().items()}, "p.", xa), atol=1e-12)


# ---- context_fn.decode: the producer of the U-Net's context list -------------------
from golden.make_golden import CTXDEC, ctxdec_latent  # noqa: E402


@pytest.mark.parametrize("case", CTXDEC, ids=[c[0] for c in CTXDEC])
def test_context_decode_matches_reference_golden(case):
    """oracle.context_decode (restatement of BigCompressor.decode / ResnetCompressor.decode) vs the unmodified
    reference's outputs on a seeded integer latent; the state_dict comes from the drop-in modules, whose keys mirror the
    reference's (so this also pins the `context_fn.dec.*` weight ABI the engine will consume for that row)."""
    from conftest import build_dropin
    name, variant, B, H, W, seed = case
    gold = np.load(os.path.join(GOLD, f"ctxdec_{name}.npz"))
    d = build_dropin(variant)
    sd = O.seeded_fill(d.state_dict(), seed=seed, denoiser_gain=0.5)
    outs = O.context_decode(sd, "context_fn.", ctxdec_latent(sd, B, H, W, seed))
    assert len(outs) == 4
    for i, o in enumerate(outs):
        ref = torch.from_numpy(gold[f"out{i}"])
        assert o.shape == ref.shape
        assert (o - ref).norm() / ref.norm() < 2e-6
    # the finest map is what the U-Net concatenates at level 0: 3 channels (eps) / 64 channels (x)
    assert outs[0].shape[1] == (3 if variant == "eps" else 64) and outs[0].shape[-2:] == (H, W)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("variant", ["eps", "x"])
def test_context_decode_matches_live_reference(variant):
    from oracle.ref_loader import build_reference_diffusion
    _, diff = build_reference_diffusion(variant, with_context_fn=True)
    sd = O.seeded_fill(diff.state_dict(), seed=4, denoiser_gain=0.5)
    diff.load_state_dict(sd)
    q = ctxdec_latent(sd, 2, 64, 32, 4)
    ref = diff.context_fn.decode(q) if variant == "x" else diff.context_fn.decode(q, None)
    for a, b in zip(O.context_decode(sd, "context_fn.", q), ref):
        assert torch.equal(a, b)
