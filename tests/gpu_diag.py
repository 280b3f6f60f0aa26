"""GPU bring-up diagnostic (run under gpurun): per-op comparison of the engine against the CPU
rounding-point emulator (tests/numerics_model.py) and the fp64 oracle.

    python tests/gpu_diag.py --variant eps --B 2 --H 64 --W 96 [--out gpurun_out/diag_eps.txt]

Prints, for every plan op with a debug view, rel-L2 / max-abs error against the emulator tap of
the same name; then the end-to-end errors vs emulator and vs oracle; then a DDIM-step check.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numerics_model as NM  # noqa: E402
from oracle import cdc_oracle as O  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="eps")
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--H", type=int, default=64)
    ap.add_argument("--W", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-taps", action="store_true")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--mainloop", type=int, default=0)
    args = ap.parse_args()
    lines = []

    def say(*a):
        s = " ".join(str(x) for x in a)
        print(s, flush=True)
        lines.append(s)

    from cdc_compression_b200 import DenoiserEngine

    variant, B, H, W = args.variant, args.B, args.H, args.W
    cc = 3 if variant == "eps" else 64
    dev = torch.device("cuda", 0)
    sd = O.seeded_unet_state_dict(variant, args.seed)
    eng = DenoiserEngine(variant, 64, (1, 2, 3, 4, 5, 6), (1, 2, 3, 4), 3, cc, dev)
    eng.load_weights(sd)
    eng.set_mainloop(args.mainloop)
    eng.set_debug(not args.no_taps)
    say(f"mainloop={args.mainloop} tc_ops={eng.tc_ops(B, H, W)}")
    ctx = O.seeded_context(variant, B, H, W, seed=args.seed)
    g = torch.Generator().manual_seed(5 + args.seed)
    x = torch.randn(B, 3, H, W, generator=g)
    t = torch.linspace(0.15, 0.9, B)
    say(f"== {variant} B={B} H={H} W={W} ops={eng.launches_per_forward(B, H, W)} "
        f"ws={eng.workspace_bytes(B, H, W) / 1e6:.1f}MB")

    t0 = time.time()
    y = eng.forward(x.to(dev), t.to(dev), [c.to(dev) for c in ctx])
    torch.cuda.synchronize()
    say(f"forward ok in {time.time() - t0:.3f}s  finite={bool(torch.isfinite(y).all())}")
    y = y.cpu()

    sd64 = {k: v.double() for k, v in sd.items()}
    ctx64 = [c.double() for c in ctx]
    with torch.no_grad():
        y64 = O.unet_forward(sd64, x.double(), t.double()[:, None], ctx64)
        NM.TAPS = {}
        yem = NM.unet_forward_emulated(sd64, x.double(), t.double()[:, None], ctx64)
        taps = NM.TAPS
        NM.TAPS = None
    say(f"END-TO-END  vs emulator rel={rel(y, yem):.3e}   vs fp64 oracle rel={rel(y, y64):.3e}   "
        f"(emulator vs oracle rel={rel(yem, y64):.3e})  maxabs_vs_oracle={(y.double() - y64).abs().max().item():.3e}")

    if not args.no_taps:
        names = eng.debug_ops(B, H, W)
        for i, name in enumerate(names):
            if name not in taps:
                continue
            got = eng.debug_read(i)
            if got is None:
                continue
            ref = taps[name].permute(0, 2, 3, 1).float()
            if got.shape != ref.shape:
                say(f"{i:4d} {name:28s} SHAPE MISMATCH got {tuple(got.shape)} ref {tuple(ref.shape)}")
                continue
            say(f"{i:4d} {name:28s} rel={rel(got, ref):.3e} maxabs={(got - ref).abs().max().item():.3e} "
                f"refmax={ref.abs().max().item():.2f} finite={bool(torch.isfinite(got).all())}")
        # packed input
        got = eng.debug_read(names.index("pack_input"))
        exp = torch.zeros(B, H, W, 64)
        src = x if variant == "x" else torch.cat([x, ctx[0]], dim=1)
        nc = src.shape[1]
        for kx in range(7):
            lo, hi = max(0, 3 - kx), min(W, W + 3 - kx)
            exp[:, :, lo:hi, kx * 8:kx * 8 + nc] = src[:, :, :, lo + kx - 3:hi + kx - 3].permute(0, 2, 3, 1)
        exp[:, :, :, 56:56 + nc] = src.permute(0, 2, 3, 1)   # slot 7 = centre copy
        say(f"pack_input maxabs={(got - exp.half().float()).abs().max().item():.3e}")

    # ---- DDIM steps: engine (eager per-step API and graph loop) vs oracle driven by the oracle U-Net ----
    S = args.steps
    var_sched = "linear" if variant == "eps" else "cosine"
    T = 20000 if variant == "eps" else 8193
    sch = O.make_sample_schedule(O.train_alphas_cumprod(var_sched, T), S, variant)
    coefs = torch.zeros(S, 8)
    for i in range(S):
        dirc = sch.one_minus_alphas_cumprod_prev[i]
        if variant == "x":
            dirc = dirc.clamp(min=0)
        coefs[i] = torch.tensor([sch.sqrt_recip_alphas_cumprod[i], sch.sqrt_recipm1_alphas_cumprod[i],
                                 sch.sqrt_alphas_cumprod_prev[i], torch.sqrt(dirc), 0.0,
                                 O.unet_time(sch, i, variant, 1).item(), torch.sqrt(sch.alphas_cumprod[i]),
                                 torch.sqrt(1 - sch.alphas_cumprod[i])])
    eng.set_debug(False)
    eng.set_schedule(coefs)
    eng.set_context([c.to(dev) for c in ctx], B, H, W)
    init = torch.randn(B, 3, H, W, generator=g) * 0.8
    pred = "noise" if variant == "eps" else "x"
    clip = "none" if variant == "eps" else "full"
    xe = init.clone().to(dev)
    for i in reversed(range(S)):
        eng.ddim_step(xe, i, None, pred, clip)
    xg = init.clone().to(dev)
    eng.sample_loop(xg, S - 1, 0, pred, clip)
    torch.cuda.synchronize()
    with torch.no_grad():
        xo = O.sample_loop(sd, sch, variant, ctx, init.clone())
    say(f"DDIM {S} steps: eager-vs-oracle rel={rel(xe.cpu(), xo):.3e}  graph-vs-eager maxabs="
        f"{(xg - xe).abs().max().item():.3e}  |x|max={xo.abs().max().item():.2f}")
    # teacher-forced single update: engine step vs oracle update using the ENGINE's own U-Net output
    i = S - 1
    xin = init.clone().to(dev)
    tt = torch.full((B,), coefs[i, 5].item(), device=dev)
    f = eng.forward(xin, tt, [c.to(dev) for c in ctx]).cpu()
    eng.set_context([c.to(dev) for c in ctx], B, H, W)
    x1 = eng.ddim_step(init.clone().to(dev), i, None, pred, clip).cpu()
    if variant == "eps":
        x1o = O.ddim_update_eps(sch, i, init, f, clip="none")
    else:
        x1o = O.ddim_update_x(sch, i, init, f, clip=True)
    say(f"DDIM teacher-forced update rel={rel(x1, x1o):.3e} maxabs={(x1 - x1o).abs().max().item():.3e}")

    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as fobj:
            fobj.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
