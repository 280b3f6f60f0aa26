"""Throughput of the other BASELINE.json configurations on one GPU (bench.py measures configs[1] only):

    python tests/gpu_configs.py [--out profiles/configs_r01.json]

config 3: xparam, B=16, 256x256, S=250          config 4 (per-GPU shard): eps, B=8, 512x512
config 5: eps, B=32, 256x256, S in {50,100,250,500} (time per step is schedule-length independent; S=50 is run)
Each line: image-steps/s, Mpix-steps/s, ms/step and the fraction of the sustained tensor roofline.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_dropin  # noqa: E402
from oracle import cdc_oracle as O  # noqa: E402


def run(variant, B, H, W, S, steps):
    dev = torch.device("cuda", 0)
    d = build_dropin(variant, with_context_fn=False)
    d.denoise_fn.load_state_dict(O.seeded_unet_state_dict(variant, 0, gain=0.5))
    d.to(dev)
    ctx = [c.to(dev) for c in O.seeded_context(variant, B, H, W)]
    x = (torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(1)) * 0.8).to(dev)
    d.set_sample_schedule(S, dev)
    eng = d._bind(x, ctx, 0.0)
    eng.set_context(ctx, B, H, W)
    pred, clip = ("noise", "none") if variant == "eps" else ("x", "full")
    eng.sample_loop(x, S - 1, S - 4, pred, clip)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.sample_loop(x, S - 5, S - 4 - steps, pred, clip)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = eng.flops_per_forward(B, H, W)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        peak = 1400.0
    return {"variant": variant, "B": B, "H": H, "W": W, "schedule": S, "timed_steps": steps, "ms_per_step": ms,
            "image_steps_per_s": B / ms * 1e3, "mpix_steps_per_s": B * H * W / ms / 1e3,
            "tflops": flops / ms / 1e9, "frac_of_sustained_tensor_peak": flops / ms / 1e9 / peak,
            "finite": bool(torch.isfinite(x).all()), "workspace_mb": eng.workspace_bytes(B, H, W) / 1e6}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = [run("x", 16, 256, 256, 250, 40), run("eps", 8, 512, 512, 500, 20), run("eps", 32, 256, 256, 50, 30),
            run("eps", 1, 256, 256, 500, 60), run("eps", 1, 512, 768, 200, 30)]
    for r in rows:
        print(json.dumps(r))
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)
