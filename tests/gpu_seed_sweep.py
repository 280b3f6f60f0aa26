"""rel-L2 of the engine's U-Net forward vs the fp64 oracle over several weight/input seeds and shapes (margin check)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cdc_oracle as O  # noqa: E402
from cdc_compression_b200 import DenoiserEngine  # noqa: E402

dev = torch.device("cuda", 0)
rows = []
for variant in ("eps", "x"):
    for seed in range(6):
        for (B, H, W) in ((1, 64, 64), (2, 32, 96), (1, 128, 128)):
            if (H, W) == (128, 128) and seed > 1:
                continue
            sd = O.seeded_unet_state_dict(variant, seed)
            eng = DenoiserEngine(variant, 64, (1, 2, 3, 4, 5, 6), (1, 2, 3, 4), 3, 3 if variant == "eps" else 64, dev)
            eng.load_weights(sd)
            ctx = O.seeded_context(variant, B, H, W, seed=seed)
            g = torch.Generator().manual_seed(50 + seed)
            x = torch.randn(B, 3, H, W, generator=g)
            t = torch.rand(B, generator=g)
            y = eng.forward(x.to(dev), t.to(dev), [c.to(dev) for c in ctx]).cpu().double()
            sd64 = {k: v.double() for k, v in sd.items()}
            with torch.no_grad():
                y64 = O.unet_forward(sd64, x.double(), t.double()[:, None], [c.double() for c in ctx])
            r = ((y - y64).norm() / y64.norm()).item()
            rows.append({"variant": variant, "seed": seed, "shape": [B, H, W], "rel_l2": r})
            print(rows[-1], flush=True)
            eng.close()
print("MAX", max(r["rel_l2"] for r in rows))
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=0)
