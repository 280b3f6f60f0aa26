"""Golden decodes of the three Kodak fixtures by the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_images.py

* x variant (BASELINE config 3 shape, shortened): each full 512x768 image, B=1, S=6, compress() -> x_hat, bpp
* eps variant (config 2's batch): eight 256x256 crops, S=6, clip_noise="full" (random-init eps models diverge
  without the clamp), compress() -> x_hat, bpp
All weights (denoiser AND context network) come from oracle.seeded_fill, so the GPU box can rebuild them.
Outputs are stored as fp16 (values in [-1,1]); PSNR(x_hat_ref, x) is stored in fp64.
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cdc_oracle as O  # noqa: E402
from oracle.ref_loader import build_reference_diffusion  # noqa: E402

S = 6


def load_images():
    out = []
    for i in (1, 2, 3):
        a = np.asarray(Image.open(os.path.join(HERE, "imgs", f"{i}.png")).convert("RGB"), dtype=np.float32) / 255.0
        out.append(torch.from_numpy(a).permute(2, 0, 1) * 2 - 1)
    return out


def init_noise(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * 0.8


def main():
    torch.set_grad_enabled(False)
    imgs = load_images()
    # ---- x variant, full images ----
    _, diff = build_reference_diffusion("x")
    diff.load_state_dict(O.seeded_fill(diff.state_dict(), seed=0, denoiser_gain=0.5))
    outs, bpps, psnrs = [], [], []
    for i, img in enumerate(imgs):
        x = img[None]
        xh, bpp = diff.compress(x, sample_steps=S, bpp_return_mean=True, init=init_noise(x.shape, 100 + i))
        outs.append(xh[0].clamp(-1, 1).half().numpy())
        bpps.append(float(bpp))
        psnrs.append(float(O.batch_psnr(xh.clamp(-1, 1) / 2 + 0.5, x / 2 + 0.5)[0]))
        print("x img", i + 1, "bpp", bpps[-1], "psnr", psnrs[-1])
    np.savez_compressed(os.path.join(HERE, "decode_x_kodak.npz"), out=np.stack(outs), bpp=np.array(bpps),
                        psnr=np.array(psnrs), S=S)
    # ---- eps variant, 8 crops ----
    _, diff = build_reference_diffusion("eps")
    diff.clip_noise = "full"
    diff.load_state_dict(O.seeded_fill(diff.state_dict(), seed=0, denoiser_gain=0.5))
    x = O.kodak_crops(imgs, 256, 8)
    xh, bpp = diff.compress(x, sample_steps=S, sample_mode="ddim", bpp_return_mean=False, init=init_noise(x.shape, 200))
    psnr = O.batch_psnr(xh.clamp(-1, 1) / 2 + 0.5, x / 2 + 0.5)
    print("eps crops bpp", bpp, "psnr", psnr)
    np.savez_compressed(os.path.join(HERE, "decode_eps_crops.npz"), out=xh.clamp(-1, 1).half().numpy(),
                        bpp=bpp.numpy(), psnr=psnr.double().numpy(), S=S)


if __name__ == "__main__":
    main()
