"""BASELINE.json configs[4]: step-count sweep 50 / 100 / 250 / 500 at batch = 32, 256 x 256 (eps), roofline fraction per block.

    python tests/gpu_config5_sweep.py [--out profiles/config5_r02.json] [--B 32]

For every schedule length S the WHOLE S-step DDIM loop is timed (CUDA events around one cdc_sample_loop call: first step
eager, the rest CUDA-graph replays); the per-block table comes from cdc_engine_profile_ops (every launch timed alone) grouped
by the rows of SURVEY.md Appendix A: algorithmic FLOPs / time against the measured burst tensor peak, and for the rows the
survey grades on HBM the algorithmic bytes / time against the measured HBM peak.
"""
import argparse
import json
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_dropin  # noqa: E402
from oracle import cdc_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--out", default=None)
a = ap.parse_args()

try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    TF, TFS, HBM, src = pk["bf16_tflops"], pk["bf16_tflops_sustained"], pk["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
except Exception:
    TF, TFS, HBM, src = 1667.0, 1386.7, 6468.3, "fallback"

dev = torch.device("cuda", 0)
B, H, W = a.B, 256, 256
d = build_dropin("eps", with_context_fn=False)
d.denoise_fn.load_state_dict(O.seeded_unet_state_dict("eps", 0, gain=0.5))
d.to(dev)
ctx = [c.to(dev) for c in O.seeded_context("eps", B, H, W)]
x0 = (torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(1)) * 0.8).to(dev)

sweep = []
eng = None
for S in (50, 100, 250, 500):
    d.set_sample_schedule(S, dev)
    x = x0.clone()
    eng = d._bind(x, ctx, 0.0)
    eng.set_context(ctx, B, H, W)
    if not sweep:   # warm-up: module loads, graph capture
        eng.sample_loop(x.clone(), S - 1, S - 6, "noise", "none")
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.sample_loop(x, S - 1, 0, "noise", "none")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flops = eng.flops_per_forward(B, H, W)
    sweep.append({"schedule_steps": S, "decode_ms": ms, "ms_per_step": ms / S, "image_steps_per_s": B * S / ms * 1e3,
                  "mpix_steps_per_s": B * H * W * S / ms / 1e3, "decoded_images_per_s": B / ms * 1e3,
                  "frac_of_sustained_tensor_peak": flops * S / ms / 1e9 / TFS, "finite": bool(torch.isfinite(x).all())})
    print(json.dumps(sweep[-1]))

prof = eng.profile_ops(iters=5)
blocks = {}
for name, ms, fl in prof:
    base = name.replace("#partials", "")
    m = re.match(r"(downs|ups)\.(\d)\.(\d)", base)
    if m:
        lvl = f"{m.group(1)}{m.group(2)}"
        part = {"0": "res0", "1": "res1", "2": "attn", "3": "resample"}[m.group(3)]
        key = f"{lvl}.{part}"
    elif base.startswith("mid_block1"):
        key = "mid.res1"
    elif base.startswith("mid_attn"):
        key = "mid.attn"
    elif base.startswith("mid_block2"):
        key = "mid.res2"
    else:
        key = base
    b = blocks.setdefault(key, {"ms": 0.0, "gflop": 0.0, "launches": 0})
    b["ms"] += ms
    b["gflop"] += fl / 1e9
    b["launches"] += 1
# rows SURVEY 8(d) grades on HBM: minimal fused fp16 traffic per image from Appendix A (MB @256^2, B=1)
hbm_mb = {"downs0.attn": 25.2, "downs0.resample": 10.5, "ups4.attn": 6.3, "ups4.resample": 10.5, "final_conv": 8.8}
tot = sum(b["ms"] for b in blocks.values())
for k, b in blocks.items():
    b["share_of_step"] = b["ms"] / tot
    b["tflops"] = b["gflop"] / b["ms"] if b["ms"] > 0 else 0.0
    b["tensor_frac_burst"] = b["tflops"] / TF
    if k in hbm_mb:
        b["hbm_gbs_single_fp16"] = hbm_mb[k] * B / b["ms"]
        b["hbm_frac"] = b["hbm_gbs_single_fp16"] / HBM
res = {"config": f"BASELINE.json configs[4]: eps, batch={B}, 256x256, step-count sweep; per-block = launches timed alone",
       "peaks": {"bf16_tflops_burst": TF, "bf16_tflops_sustained": TFS, "hbm_gbs": HBM, "source": src},
       "sweep": sweep, "per_block_sum_ms": tot, "blocks": blocks}
for k, b in blocks.items():
    print(f"{k:18s} {b['ms']*1e3:8.1f} us {b['share_of_step']*100:5.1f} %  {b['tflops']:7.1f} TFLOP/s  frac {b['tensor_frac_burst']:.3f}"
          + (f"  HBM {b['hbm_gbs_single_fp16']:.0f} GB/s ({b['hbm_frac']:.2f})" if "hbm_frac" in b else ""))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
