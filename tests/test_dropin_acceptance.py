"""BASELINE.json's acceptance sentence, automated: the reference's own demo scripts (epsilonparam/test_epsilonparam.py,
xparam/test_xparam.py) run UNCHANGED against the drop-in modules, with the variant directory as cwd, a synthetic
checkpoint in the reference layout ({"model": sd} incl. loss_fn_vgg.* keys for --lpips_weight 0.9 / {"ema": EMA
state_dict}), and the PNGs they save are compared with the same decode done in-process through the public API.

The scripts themselves are reference sources: they are not committed.  `__graft_entry__.build()` copies them verbatim
from /root/reference into oracle/_ref/scripts/ (git-ignored; ships to the GPU box with the snapshot)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import cdc_oracle as O
from conftest import build_dropin, import_variant

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(ROOT, "oracle", "_ref", "scripts")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 20260117   # the default CUDA generator of a fresh process is seeded from the OS (verified: two runs of the script gave
                  # different init noise), and the scripts never seed: the test seeds the child through a sitecustomize
                  # module on PYTHONPATH — the scripts themselves stay byte-for-byte the reference's
torch.set_grad_enabled(False)


def _script(name):
    path = os.path.join(SCRIPTS, name)
    if not os.path.isfile(path):
        pytest.skip(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists")
    return path


def _crops(tmp_path):
    """Two 128x192 crops of the Kodak fixtures as PNG files (the scripts loop over a directory)."""
    from PIL import Image
    d = tmp_path / "imgs"
    d.mkdir()
    for i, (y, x) in ((1, (64, 96)), (2, (256, 320))):
        im = Image.open(os.path.join(GOLD, "imgs", f"{i}.png")).convert("RGB").crop((x, y, x + 192, y + 128))
        im.save(d / f"crop{i}.png")
    return d


def _run(script, variant_dir, args):
    site = os.path.join(os.path.dirname(args[args.index("--ckpt") + 1]), "site")
    os.makedirs(site, exist_ok=True)
    with open(os.path.join(site, "sitecustomize.py"), "w") as f:
        f.write(f"import torch\ntorch.manual_seed({SEED})\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([site, variant_dir, os.environ.get("PYTHONPATH", "")]))
    res = subprocess.run([sys.executable, script] + args, cwd=variant_dir, env=env, capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    return res.stdout


def _read_png(path):
    import torchvision
    return torchvision.io.read_image(str(path))


def _expected(d, img_dir, steps, gamma, variant):
    """What the script computes, through the same public calls, consuming the CUDA generator identically."""
    import torchvision
    dev = torch.device("cuda", 0)
    torch.manual_seed(SEED)
    out = {}
    for name in os.listdir(img_dir):
        x = torchvision.io.read_image(os.path.join(img_dir, name)).unsqueeze(0).float().to(dev) / 255.0
        init = torch.randn_like(x) * gamma
        if variant == "eps":
            y, bpp = d.compress(x * 2.0 - 1.0, sample_steps=steps, sample_mode="ddim", bpp_return_mean=False, init=init)
        else:
            y, bpp = d.compress(x * 2.0 - 1.0, sample_steps=steps, bpp_return_mean=True, init=init)
        y = y.clamp(-1, 1) / 2.0 + 0.5
        out[name] = (y.cpu()[0].mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8), bpp)   # torchvision.utils.save_image
    return out


@pytest.mark.gpu
def test_reference_eps_demo_script_runs_unchanged(tmp_path):
    script = _script("test_epsilonparam.py")
    vdir = os.path.join(ROOT, "cdc_compression_b200", "epsilonparam")
    imgs, outd = _crops(tmp_path), tmp_path / "out"
    d = build_dropin("eps")
    sd = O.seeded_fill(d.state_dict(), seed=0, denoiser_gain=0.5)
    ckpt = dict(sd)
    ckpt["loss_fn_vgg.lin0.model.1.weight"] = torch.zeros(1, 64, 1, 1)   # saved by training runs with lpips_weight > 0
    torch.save({"model": ckpt}, tmp_path / "eps.pt")
    stdout = _run(script, vdir, ["--ckpt", str(tmp_path / "eps.pt"), "--lpips_weight", "0.9", "--n_denoise_step", "6",
                                 "--img_dir", str(imgs), "--out_dir", str(outd)])
    assert stdout.count("bpp:") == 2
    d.load_state_dict(sd)
    d.to(torch.device("cuda", 0))
    exp = _expected(d, str(imgs), 6, 0.8, "eps")
    if os.environ.get("CDC_ACCEPT_DIAG"):     # diagnostics: is the script reproducible run to run? is the in-process decode?
        outd2 = tmp_path / "out2"
        _run(script, vdir, ["--ckpt", str(tmp_path / "eps.pt"), "--lpips_weight", "0.9", "--n_denoise_step", "6",
                            "--img_dir", str(imgs), "--out_dir", str(outd2)])
        exp2 = _expected(d, str(imgs), 6, 0.8, "eps")
        fr = lambda a, b: ((a.int() - b.int()).abs() > 1).float().mean().item()
        for name in exp:
            print(f"\n[diag] {name} (listdir order {os.listdir(str(imgs))}): script1 vs script2 "
                  f"{fr(_read_png(outd / name), _read_png(outd2 / name)):.4f}, in-process 1 vs 2 "
                  f"{fr(exp[name][0], exp2[name][0]):.4f}, script vs in-process {fr(_read_png(outd / name), exp[name][0]):.4f}")
    for name, (png, _) in exp.items():
        got = _read_png(outd / name)
        assert got.shape == png.shape
        # Same engine, same inputs.  A random-init eps model with clip_noise="none" (hard-coded in the script) decodes to
        # |x| ~ 10^2, so the saved image is saturated (0 / 255) and a pixel whose value is ~0 flips on a 1-ulp difference
        # in the cuDNN context network between two processes: gate the FRACTION of pixels that differ by more than 1 LSB.
        frac = ((got.int() - png.int()).abs() > 1).float().mean().item()
        assert frac < 5e-3, (name, frac)


@pytest.mark.gpu
def test_reference_x_demo_script_runs_unchanged(tmp_path):
    script = _script("test_xparam.py")
    vdir = os.path.join(ROOT, "cdc_compression_b200", "xparam")
    imgs, outd = _crops(tmp_path), tmp_path / "out"
    d = build_dropin("x")
    sd = O.seeded_fill(d.state_dict(), seed=1, denoiser_gain=0.5)
    d.load_state_dict(sd)
    sys.path.insert(0, vdir)
    try:
        sys.modules.pop("ema_pytorch", None)
        from ema_pytorch import EMA
    finally:
        sys.path.remove(vdir)
    ema = EMA(d, beta=0.999, update_every=10, power=0.75, update_after_step=100)
    ema_sd = ema.state_dict()
    # the EMA weights differ from the online ones in a real checkpoint: the script must decode with ema_model.*
    for k in list(ema_sd):
        if k.startswith("online_model.denoise_fn.final_conv.1."):
            ema_sd[k] = ema_sd[k] * 0.0
    torch.save({"ema": ema_sd}, tmp_path / "x.pt")
    stdout = _run(script, vdir, ["--ckpt", str(tmp_path / "x.pt"), "--lpips_weight", "0.0", "--n_denoise_step", "7",
                                 "--img_dir", str(imgs), "--out_dir", str(outd)])
    assert stdout.count("bpp:") == 2
    d.to(torch.device("cuda", 0))
    exp = _expected(d, str(imgs), 7, 0.8, "x")
    for name, (png, _) in exp.items():
        got = _read_png(outd / name)
        assert got.shape == png.shape
        assert (got.int() - png.int()).abs().max().item() <= 1, name


def test_reference_demo_scripts_construct_and_load_on_cpu(tmp_path):
    """CPU half of the acceptance (no GPU here): both scripts, unchanged, import the drop-in modules, construct the
    model and strict-load a reference-layout checkpoint; they stop at `.to(0)` only because there is no CUDA device."""
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU tests")
    for variant, name, sub in (("eps", "test_epsilonparam.py", "epsilonparam"), ("x", "test_xparam.py", "xparam")):
        script = _script(name)
        vdir = os.path.join(ROOT, "cdc_compression_b200", sub)
        d = build_dropin(variant)
        sd = O.seeded_fill(d.state_dict(), seed=0)
        if variant == "eps":
            torch.save({"model": sd}, tmp_path / "c.pt")
        else:
            torch.save({"ema": {**{"ema_model." + k: v for k, v in sd.items()},
                                **{"online_model." + k: v for k, v in sd.items()},
                                "initted": torch.tensor([True]), "step": torch.tensor([1000])}}, tmp_path / "c.pt")
        env = dict(os.environ, PYTHONPATH=vdir + os.pathsep + os.environ.get("PYTHONPATH", ""))
        res = subprocess.run([sys.executable, script, "--ckpt", str(tmp_path / "c.pt"), "--lpips_weight", "0.0",
                              "--img_dir", os.path.join(GOLD, "imgs"), "--out_dir", str(tmp_path / "o")], cwd=vdir,
                             env=env, capture_output=True, text=True, timeout=600)
        assert res.returncode != 0
        tail = res.stderr[-1500:]
        assert "diffusion.to(rank)" in tail and ("NVIDIA" in tail or "CUDA" in tail or "cuda" in tail), tail
