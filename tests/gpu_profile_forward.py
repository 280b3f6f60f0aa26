"""Run N U-Net forwards of the bench shape (for ncu): python tests/gpu_profile_forward.py [--B 8 --H 256 --W 256 --iters 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cdc_oracle as O  # noqa: E402
from cdc_compression_b200 import DenoiserEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="eps")
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--H", type=int, default=256)
ap.add_argument("--W", type=int, default=256)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--mainloop", type=int, default=1)
ap.add_argument("--names-out", default=None, help="write the plan's op names (launch order of one forward) as JSON")
a = ap.parse_args()
dev = torch.device("cuda", 0)
eng = DenoiserEngine(a.variant, 64, (1, 2, 3, 4, 5, 6), (1, 2, 3, 4), 3, 3 if a.variant == "eps" else 64, dev)
eng.load_weights(O.seeded_unet_state_dict(a.variant, 0, gain=0.5))
eng.set_mainloop(a.mainloop)
ctx = [c.to(dev) for c in O.seeded_context(a.variant, a.B, a.H, a.W)]
x = torch.randn(a.B, 3, a.H, a.W, device=dev)
t = torch.full((a.B,), 0.5, device=dev)
for _ in range(a.iters):
    y = eng.forward(x, t, ctx)
torch.cuda.synchronize()
if a.names_out:
    import json
    json.dump(eng.debug_ops(a.B, a.H, a.W), open(a.names_out, "w"))
print("ok", float(y.abs().mean()), "tc ops", eng.tc_ops(a.B, a.H, a.W))
