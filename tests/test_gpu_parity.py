"""Parity of the CUDA path (through the C ABI and the drop-in modules) against the CPU oracle and the
golden vectors produced by the unmodified reference.  Run on the B200 box:  pytest -m gpu

Tolerances (floating point path; north_star: "within 1e-3 rel fp16"):
  * REL_L2_UNET      rel-L2 of one U-Net forward vs the reference's fp32 output.
  * teacher-forced sampler arithmetic is fp32 and must match to ~1e-6.
"""
import os

import numpy as np
import pytest
import torch

from oracle import cdc_oracle as O
import numerics_model as NM
from conftest import build_dropin
from golden.make_golden import CASES, LOOPS, case_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_L2_UNET = 1.0e-3   # north_star tolerance; measured 6.2e-4 (eps) / 6.7e-4 (x) with the compensated trunk (DESIGN.md "precision")
torch.set_grad_enabled(False)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def dev():
    assert torch.cuda.is_available(), "GPU test selected without a CUDA device"
    return torch.device("cuda", 0)


_ENGINES = {}


def unet_on_gpu(variant, seed, gain=1.0):
    key = (variant, seed, gain)
    if key not in _ENGINES:
        d = build_dropin(variant, with_context_fn=False)
        d.denoise_fn.load_state_dict(O.seeded_unet_state_dict(variant, seed, gain=gain))
        d.to(dev())
        _ENGINES[key] = d
    return _ENGINES[key]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_unet_forward_vs_reference_golden(case):
    name, variant, B, H, W, seed, gain = case
    gold = torch.from_numpy(np.load(os.path.join(GOLD, f"unet_{name}.npz"))["out"])
    d = unet_on_gpu(variant, seed, gain)
    x, t, ctx, _ = case_inputs(variant, B, H, W, seed)
    y = d.denoise_fn(x.to(dev()), t.to(dev()), [c.to(dev()) for c in ctx])
    assert y.shape == gold.shape and y.dtype == torch.float32 and y.is_cuda
    assert torch.isfinite(y).all()
    assert rel(y, gold) < REL_L2_UNET, rel(y, gold)


@pytest.mark.parametrize("variant", ["eps", "x"])
def test_unet_forward_vs_oracle_and_emulator(variant):
    """GPU error vs the fp64 oracle must be what the rounding-point emulator predicts (no hidden bug)."""
    B, H, W, seed = 2, 64, 64, 2
    d = unet_on_gpu(variant, seed)
    x, t, ctx, _ = case_inputs(variant, B, H, W, seed)
    y = d.denoise_fn(x.to(dev()), t.to(dev()), [c.to(dev()) for c in ctx])
    sd64 = {k: v.double() for k, v in O.seeded_unet_state_dict(variant, seed).items()}
    y64 = O.unet_forward(sd64, x.double(), t.double(), [c.double() for c in ctx])
    yem = NM.unet_forward_emulated(sd64, x.double(), t.double(), [c.double() for c in ctx])
    e_gpu, e_em = rel(y, y64), rel(yem, y64)
    assert e_gpu < REL_L2_UNET
    assert e_gpu < 1.3 * e_em + 1e-4, (e_gpu, e_em)


@pytest.mark.parametrize("variant", ["eps", "x"])
def test_tcgen05_and_hmma_mainloops_agree(variant):
    """The tcgen05/TMA convolution kernel (default) and the mma.sync baseline kernel implement the same rounding
    points: the first Block's output may differ only by fp32 summation order, and both meet the tolerance."""
    B, H, W, seed = 2, 64, 96, 0       # 96-wide: ragged 16-pixel tiles at every level
    d = unet_on_gpu(variant, seed)
    eng = d.denoise_fn.engine_for(dev())
    x, t, ctx, _ = case_inputs(variant, B, H, W, seed)
    ctxd = [c.to(dev()) for c in ctx]
    outs, first = {}, {}
    try:
        for kind in (0, 1):
            eng.set_mainloop(kind)
            eng.set_debug(True)
            assert (eng.tc_ops(B, H, W) > 0) == (kind == 1)
            outs[kind] = eng.forward(x.to(dev()), t.reshape(-1).to(dev()), ctxd).cpu()
            names = eng.debug_ops(B, H, W)
            first[kind] = eng.debug_read(names.index("downs.0.0.block1"))
    finally:
        eng.set_debug(False)
        eng.set_mainloop(1)
    assert rel(first[1], first[0]) < 5e-5
    sd64 = {k: v.double() for k, v in O.seeded_unet_state_dict(variant, seed).items()}
    y64 = O.unet_forward(sd64, x.double(), t.double(), [c.double() for c in ctx])
    assert rel(outs[0], y64) < REL_L2_UNET and rel(outs[1], y64) < REL_L2_UNET


@pytest.mark.parametrize("variant", ["eps", "x"])
def test_batch_elements_are_independent(variant):
    B, H, W, seed = 3, 32, 64, 0
    d = unet_on_gpu(variant, seed)
    x, _, ctx, _ = case_inputs(variant, B, H, W, seed)
    t = torch.tensor([[0.2], [0.5], [0.8]])
    yb = d.denoise_fn(x.to(dev()), t.to(dev()), [c.to(dev()) for c in ctx])
    for i in range(B):
        yi = d.denoise_fn(x[i:i + 1].to(dev()), t[i:i + 1].to(dev()), [c[i:i + 1].to(dev()) for c in ctx])
        assert torch.equal(yb[i:i + 1], yi)     # batch-invariant: no kernel's summation order depends on B


def _coefs(sch, variant, eta=0.0):
    S = sch.sample_steps
    c = torch.zeros(S, 8)
    for i in range(S):
        dirc = sch.one_minus_alphas_cumprod_prev[i] - (eta * sch.sigma[i]) ** 2
        if variant == "x":
            dirc = dirc.clamp(min=0)
        c[i] = torch.tensor([sch.sqrt_recip_alphas_cumprod[i], sch.sqrt_recipm1_alphas_cumprod[i],
                             sch.sqrt_alphas_cumprod_prev[i], torch.sqrt(dirc), eta * sch.sigma[i],
                             O.unet_time(sch, i, variant, 1).item(), torch.sqrt(sch.alphas_cumprod[i]),
                             torch.sqrt(1 - sch.alphas_cumprod[i])])
    return c


@pytest.mark.parametrize("variant,clip,eta", [("eps", "none", 0.0), ("eps", "full", 0.0), ("eps", "half", 0.0),
                                              ("x", "full", 0.0), ("x", "none", 0.0), ("x", "full", 0.5),
                                              ("eps", "none", 0.7)])
def test_ddim_update_teacher_forced(variant, clip, eta):
    """One engine DDIM step vs the oracle's update fed with the ENGINE's own U-Net output: isolates the fused
    eps -> x0 -> x_{t-1} arithmetic (fp32) from the fp16 U-Net."""
    B, H, W, seed, S = 2, 32, 32, 0, 7
    d = unet_on_gpu(variant, seed)
    eng = d.denoise_fn.engine_for(dev())
    T, sched = (20000, "linear") if variant == "eps" else (8193, "cosine")
    sch = O.make_sample_schedule(O.train_alphas_cumprod(sched, T), S, variant)
    coefs = _coefs(sch, variant, eta)
    _, _, ctx, init = case_inputs(variant, B, H, W, seed)
    ctxd = [c.to(dev()) for c in ctx]
    g = torch.Generator().manual_seed(9)
    z = torch.randn(init.shape, generator=g)
    pred = "noise" if variant == "eps" else "x"
    for i in (S - 1, 3, 0):
        tt = torch.full((B,), coefs[i, 5].item(), device=dev())
        f = eng.forward(init.to(dev()), tt, ctxd).cpu()
        eng.set_schedule(coefs)
        eng.set_context(ctxd, B, H, W)
        x1 = eng.ddim_step(init.clone().to(dev()), i, z.to(dev()) if eta else None, pred, clip).cpu()
        if variant == "eps":
            ref = O.ddim_update_eps(sch, i, init, f, clip=clip, eta=eta, z=z)
        else:
            ref = O.ddim_update_x(sch, i, init, f, clip=(clip == "full"), eta=eta, z=z)
        assert torch.allclose(x1, ref, rtol=2e-5, atol=2e-5 * ref.abs().max().item()), (i, (x1 - ref).abs().max())


@pytest.mark.parametrize("variant", ["eps", "x"])
def test_graph_loop_equals_eager_steps(variant):
    B, H, W, seed, S = 2, 32, 32, 0, 5
    d = unet_on_gpu(variant, seed)
    eng = d.denoise_fn.engine_for(dev())
    T, sched = (20000, "linear") if variant == "eps" else (8193, "cosine")
    sch = O.make_sample_schedule(O.train_alphas_cumprod(sched, T), S, variant)
    eng.set_schedule(_coefs(sch, variant))
    _, _, ctx, init = case_inputs(variant, B, H, W, seed)
    eng.set_context([c.to(dev()) for c in ctx], B, H, W)
    pred, clip = ("noise", "none") if variant == "eps" else ("x", "full")
    xe = init.clone().to(dev())
    for i in reversed(range(S)):
        eng.ddim_step(xe, i, None, pred, clip)
    for _ in range(2):   # first call captures the graph, second replays it
        xg = init.clone().to(dev())
        eng.sample_loop(xg, S - 1, 0, pred, clip)
        assert torch.equal(xg, xe)


@pytest.mark.parametrize("case", LOOPS, ids=[c[0] for c in LOOPS])
def test_sample_loop_vs_reference_golden(case):
    name, variant, B, H, W, S, seed = case
    gold = torch.from_numpy(np.load(os.path.join(GOLD, f"loop_{name}.npz"))["out"])
    d = unet_on_gpu(variant, seed)
    _, _, ctx, init = case_inputs(variant, B, H, W, seed)
    d.set_sample_schedule(S, dev())
    ctxd = [c.to(dev()) for c in ctx]
    if variant == "eps":
        out = d.p_sample_loop(init.shape, ctxd, "ddim", init=init.to(dev()), eta=0)
        # random-init eps trajectories amplify every per-step error by sqrt(1/acp) up to ~150x
        assert rel(out, gold) < 0.05
    else:
        out = d.p_sample_loop(init.shape, ctxd, clip_denoised=True, init=init.to(dev()), eta=0)
        psnr = O.batch_psnr(out.cpu().clamp(-1, 1) / 2 + 0.5, gold.clamp(-1, 1) / 2 + 0.5)
        assert psnr.min() > 45.0, psnr


@pytest.mark.parametrize("variant", ["eps", "x"])
def test_compress_dropin_end_to_end(variant):
    """diffusion.compress() through the drop-in modules (context_fn in PyTorch on the GPU, denoiser + DDIM on the
    engine) vs the same context fed to the CPU oracle loop."""
    torch.manual_seed(0)
    d = build_dropin(variant)
    sd = d.state_dict()
    for k, v in O.seeded_unet_state_dict(variant, 4, gain=0.5).items():
        sd["denoise_fn." + k] = v
    d.load_state_dict(sd)
    d.to(dev())
    g = torch.Generator().manual_seed(3)
    img = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    init = torch.randn(2, 3, 64, 64, generator=g) * 0.8
    S = 6
    if variant == "eps":
        out, bpp = d.compress(img.to(dev()), sample_steps=S, sample_mode="ddim", bpp_return_mean=False,
                              init=init.to(dev()))
        ctx = d.context_fn(img.to(dev()), None)["output"]
    else:
        out, bpp = d.compress(img.to(dev()), sample_steps=S, bpp_return_mean=True, init=init.to(dev()))
        ctx = d.context_fn(img.to(dev()))["output"]
    assert out.shape == img.shape and torch.isfinite(out).all() and torch.isfinite(bpp).all()
    T, sched = (20000, "linear") if variant == "eps" else (8193, "cosine")
    sch = O.make_sample_schedule(O.train_alphas_cumprod(sched, T), S, variant)
    ref = O.sample_loop(O.sub_state_dict({k: v.cpu() for k, v in d.state_dict().items()}, "denoise_fn."), sch, variant,
                        [c.cpu() for c in ctx], init.clone())
    to01 = lambda v: v.cpu().clamp(-1, 1) / 2 + 0.5
    psnr = O.batch_psnr(to01(out), to01(ref))
    assert psnr.min() > 40.0, psnr


def _kodak():
    from PIL import Image
    out = []
    for i in (1, 2, 3):
        a = np.asarray(Image.open(os.path.join(GOLD, "imgs", f"{i}.png")).convert("RGB"), dtype=np.float32) / 255.0
        out.append(torch.from_numpy(a).permute(2, 0, 1) * 2 - 1)
    return out


def _init_noise(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * 0.8


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False          # context_fn stays PyTorch: keep its convs in true fp32
    torch.backends.cuda.matmul.allow_tf32 = False


def test_kodak_decode_x_variant_matches_reference_psnr():
    """BASELINE config 3 (shortened to S=6): the three 512x768 Kodak fixtures through the drop-in x-variant
    compress() vs the decode produced by the UNMODIFIED reference (tests/golden/make_golden_images.py).
    Gate: |PSNR(ours, x) - PSNR(ref, x)| <= 0.01 dB.  LPIPS cannot be measured offline (no VGG weights)."""
    _no_tf32()
    gold = np.load(os.path.join(GOLD, "decode_x_kodak.npz"))
    d = build_dropin("x")
    d.load_state_dict(O.seeded_fill(d.state_dict(), seed=0, denoiser_gain=0.5))
    d.to(dev())
    for i, img in enumerate(_kodak()):
        x = img[None]
        xh, bpp = d.compress(x.to(dev()), sample_steps=int(gold["S"]), bpp_return_mean=True,
                             init=_init_noise(x.shape, 100 + i).to(dev()))
        xh = xh.cpu().clamp(-1, 1)
        ref = torch.from_numpy(gold["out"][i]).float()[None]
        psnr_ours = O.batch_psnr(xh / 2 + 0.5, x / 2 + 0.5)[0].item()
        assert abs(psnr_ours - float(gold["psnr"][i])) <= 0.01, (i, psnr_ours, float(gold["psnr"][i]))
        assert O.batch_psnr(xh / 2 + 0.5, ref / 2 + 0.5)[0].item() > 45.0
        assert abs(float(bpp) - float(gold["bpp"][i])) < 1e-3 * float(gold["bpp"][i])


def test_kodak_crops_decode_eps_variant_matches_reference_psnr():
    """BASELINE config 2's batch (eight 256x256 Kodak crops) through the drop-in eps-variant compress(), S=6,
    clip_noise="full" (random-init eps models diverge without the clamp), vs the unmodified reference's decode."""
    _no_tf32()
    gold = np.load(os.path.join(GOLD, "decode_eps_crops.npz"))
    d = build_dropin("eps")
    d.clip_noise = "full"
    d.load_state_dict(O.seeded_fill(d.state_dict(), seed=0, denoiser_gain=0.5))
    d.to(dev())
    x = O.kodak_crops(_kodak(), 256, 8)
    xh, bpp = d.compress(x.to(dev()), sample_steps=int(gold["S"]), sample_mode="ddim", bpp_return_mean=False,
                         init=_init_noise(x.shape, 200).to(dev()))
    xh = xh.cpu().clamp(-1, 1)
    ref = torch.from_numpy(gold["out"]).float()
    psnr_ours = O.batch_psnr(xh / 2 + 0.5, x / 2 + 0.5)
    assert (psnr_ours - torch.from_numpy(gold["psnr"]).float()).abs().max().item() <= 0.01
    assert O.batch_psnr(xh / 2 + 0.5, ref / 2 + 0.5).min().item() > 40.0
    assert torch.allclose(bpp.cpu(), torch.from_numpy(gold["bpp"]).float(), rtol=1e-3)


def test_rng_stream_advances_like_reference():
    """The reference draws randn_like once per step even at eta=0; the drop-in leaves the CUDA generator in
    the same state so the next image's init noise is identical."""
    d = unet_on_gpu("x", 0)
    _, _, ctx, init = case_inputs("x", 1, 32, 32, 0)
    S = 4
    d.set_sample_schedule(S, dev())
    torch.cuda.manual_seed(123)
    d.p_sample_loop(init.shape, [c.to(dev()) for c in ctx], clip_denoised=True, init=init.to(dev()))
    a = torch.randn(1, 3, 32, 32, device=dev())
    torch.cuda.manual_seed(123)
    for _ in range(S):
        torch.randn_like(init.to(dev()))
    b = torch.randn(1, 3, 32, 32, device=dev())
    assert torch.equal(a, b)


def test_errors_are_loud():
    from cdc_compression_b200 import EngineError
    d = unet_on_gpu("eps", 0)
    x, t, ctx, _ = case_inputs("eps", 1, 32, 32, 0)
    with pytest.raises(EngineError):      # CPU tensors: no CPU path
        d.denoise_fn(x, t, ctx)
    with pytest.raises(EngineError):      # H not a multiple of 32
        d.denoise_fn(torch.zeros(1, 3, 48, 32, device=dev()), t.to(dev()), [c.to(dev()) for c in ctx])
    with pytest.raises(EngineError):      # wrong number of context maps
        d.denoise_fn(x.to(dev()), t.to(dev()), [c.to(dev()) for c in ctx[:2]])
    eng = d.denoise_fn.engine_for(dev())
    with pytest.raises(EngineError):      # sampler before set_context for this shape
        eng.ddim_step(torch.zeros(1, 3, 96, 96, device=dev()), 0, None, "noise", "none")


def test_large_shape_512_runs_and_is_finite():
    """BASELINE config 4's per-GPU shape (8 x 512 x 512 is exercised by bench.py); here one 512x768 Kodak-sized
    image — the demo scripts' actual input — checked through a size-independent property (finite, and equal to
    the same image decoded inside a batch of two)."""
    d = unet_on_gpu("eps", 0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 512, 768, generator=g)
    ctx = O.seeded_context("eps", 2, 512, 768)
    t = torch.tensor([[0.4], [0.4]])
    y2 = d.denoise_fn(x.to(dev()), t.to(dev()), [c.to(dev()) for c in ctx])
    y1 = d.denoise_fn(x[:1].to(dev()), t[:1].to(dev()), [c[:1].to(dev()) for c in ctx])
    assert torch.isfinite(y2).all()
    assert torch.equal(y2[:1], y1)


KERNEL_FORMS = [("CDC_ATTN_TC", "mma.sync attention-context kernel instead of the tcgen05 one"),
                ("CDC_NSLICE", "K-split partial tiles + ln_rows_kernel instead of fused column slices / cluster LayerNorm"),
                ("CDC_FINAL_KX", "49-tap final convolution instead of the horizontal-taps-in-N form"),
                ("CDC_SLICED", "whole-row tiles at every level"),
                ("CDC_FOLD_FINISH", "separate attn_finish_kernel instead of the fused finish epilogue"),
                ("CDC_FINAL_TC", "mma.sync final convolution instead of the tcgen05 one"),
                ("CDC_DUAL_PASS", "separate W_hi / W_lo passes instead of one activation load feeding two weight tiles"),
                ("CDC_FUSE_RES", "separate res_conv launches instead of the second TMEM accumulator in block2"),
                ("CDC_FINAL_PRELN", "final convolution normalises its own halo instead of reading the last Upsample's LayerNorm-ed copy"),
                ("CDC_FINAL_PERSIST", "persistent double-buffered final convolution vs one 16 x 16 tile per CTA"),
                ("CDC_ALG_TC", "attention C x C products on tcgen05 (MN-major operands) vs the split-fp16 mma.sync kernel"),
                ("CDC_PDL", "programmatic dependent launch on every kernel (default) vs plain stream order"),
                ("CDC_FUSE_LNROWS", "K-split convolutions finish their rows in the launch (arrival counters) vs the separate ln_rows_kernel")]


@pytest.mark.parametrize("knob", [k for k, _ in KERNEL_FORMS])
def test_alternative_kernel_forms_agree(knob):
    """Every optimised kernel form has a simpler sibling selected by an environment knob (read when an engine is
    created).  Both forms implement the same rounding points; their outputs differ by fp32 summation order (and, for
    the attention kernel, by the softmax reference maximum) — inside the oracle tolerance.
    128x128 / B=2 reaches: cluster LayerNorm at levels 1-3, the legacy sliced form below, attention tc kernel at
    C = 64..320, single-chunk context normalisation at C = 384."""
    from cdc_compression_b200 import DenoiserEngine
    variant, B, H, W, seed = "eps", 2, 128, 128, 3
    sd = O.seeded_unet_state_dict(variant, seed)
    x, t, ctx, _ = case_inputs(variant, B, H, W, seed)
    xd, td, ctxd = x.to(dev()), t.reshape(-1).to(dev()), [c.to(dev()) for c in ctx]
    outs = {}
    old = os.environ.get(knob)
    try:
        for val in ("1", "0"):
            os.environ[knob] = val
            eng = DenoiserEngine(variant, 64, (1, 2, 3, 4, 5, 6), (1, 2, 3, 4), 3, 3, dev())
            eng.load_weights(sd)
            outs[val] = eng.forward(xd, td, ctxd).cpu()
            del eng
    finally:
        if old is None:
            os.environ.pop(knob, None)
        else:
            os.environ[knob] = old
    sd64 = {k: v.double() for k, v in sd.items()}
    y64 = O.unet_forward(sd64, x.double(), t.double(), [c.double() for c in ctx])
    assert rel(outs["1"], y64) < REL_L2_UNET and rel(outs["0"], y64) < REL_L2_UNET
    # A different fp32 summation order flips a few fp16 roundings early in the network and the difference grows to
    # the size of the rounding noise itself (measured 4.0e-4 .. 4.8e-4 between forms, each 6e-4 from the oracle):
    # the forms are two draws of the same noise, so they must agree to within the oracle tolerance, not better.
    tol = 8e-4
    assert rel(outs["1"], outs["0"]) < tol, rel(outs["1"], outs["0"])
