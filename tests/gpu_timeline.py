"""In-graph kernel timeline of one denoising step (CUPTI through torch.profiler; no serialisation, PDL overlap visible):

    python tests/gpu_timeline.py [--B 8 --H 256 --W 256 --out gpurun_out/timeline.json]

For every kernel of one CUDA-graph replay: op name (plan order), start offset, duration, gap to the end of the
previous kernel on the critical path.  cdc_engine_profile_ops times launches ALONE; this shows what the graph does.
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

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="eps")
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--H", type=int, default=256)
ap.add_argument("--W", type=int, default=256)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--out", default=None)
a = ap.parse_args()

dev = torch.device("cuda", 0)
d = build_dropin(a.variant, with_context_fn=False)
d.denoise_fn.load_state_dict(O.seeded_unet_state_dict(a.variant, 0, gain=0.5))
d.to(dev)
ctx = [c.to(dev) for c in O.seeded_context(a.variant, a.B, a.H, a.W)]
x = (torch.randn(a.B, 3, a.H, a.W, generator=torch.Generator().manual_seed(1)) * 0.8).to(dev)
S = 500
d.set_sample_schedule(S, dev)
eng = d._bind(x, ctx, 0.0)
eng.set_context(ctx, a.B, a.H, a.W)
pred, clip = ("noise", "none") if a.variant == "eps" else ("x", "full")
eng.sample_loop(x, S - 1, S - 8, pred, clip)
torch.cuda.synchronize()
names = eng.debug_ops(a.B, a.H, a.W)

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.sample_loop(x, S - 9, S - 8 - a.steps, pred, clip)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name
      and "Memset" not in e.name]
ev.sort(key=lambda e: e.time_range.start)
per_step = len(ev) // a.steps
print("kernels", len(ev), "per step", per_step, "plan ops", len(names))
# take the second-to-last full step
k0 = per_step * (a.steps - 2)
step = ev[k0:k0 + per_step]
t0 = step[0].time_range.start
rows = []
prev_end = t0
for i, e in enumerate(step):
    s, en = e.time_range.start - t0, e.time_range.end - t0
    rows.append({"i": i, "kernel": e.name.split("<")[0].split("(")[0][-40:], "start_us": s, "dur_us": en - s,
                 "gap_us": e.time_range.start - prev_end})
    prev_end = max(prev_end, e.time_range.end)
total = prev_end - t0
nxt = ev[k0 + per_step].time_range.start - t0 if k0 + per_step < len(ev) else None
print("step span us", total, "next step starts at", nxt)
for r in rows:
    print(f"{r['i']:4d} {r['kernel']:40s} start {r['start_us']:9.1f} dur {r['dur_us']:7.1f} gap {r['gap_us']:7.1f}")
if a.out:
    json.dump({"names": names, "rows": rows, "span_us": total, "next_start_us": nxt}, open(a.out, "w"), indent=0)
