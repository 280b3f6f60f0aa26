"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header declares; the
planning-only engine validates weights, sizes workspaces and counts FLOPs; the product path refuses to run
without CUDA; the batch-sharding helper works across 2 gloo ranks."""
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import cdc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib_built):
    header = open(os.path.join(ROOT, "include", "cdc_b200.h")).read()
    declared = set(re.findall(r"\b(cdc_[a-z_0-9]+)\s*\(", header))
    from cdc_compression_b200 import _native
    assert declared == set(_native.exported_symbols()), declared ^ set(_native.exported_symbols())
    out = subprocess.run(["nm", "-D", "--defined-only", lib_built], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (cdc_[a-z_0-9]+)", out))
    assert declared <= exported, declared - exported
    lib = _native.load()
    assert lib.cdc_abi_version() == _native.CDC_ABI_VERSION


def test_sass_is_sm100a(lib_built):
    out = subprocess.run(["cuobjdump", "--list-elf", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.fixture(scope="module")
def planner(lib_built):
    from cdc_compression_b200 import DenoiserEngine
    engines = {}
    for variant, cc in (("eps", 3), ("x", 64)):
        e = DenoiserEngine(variant, 64, (1, 2, 3, 4, 5, 6), (1, 2, 3, 4), 3, cc, None)
        e.load_weights(O.seeded_unet_state_dict(variant, 0))
        engines[variant] = e
    return engines


def test_algorithmic_flops_match_survey(planner):
    # FLOPs per 256x256 image-step, exactly linear in B*H*W
    assert planner["eps"].flops_per_forward(1, 256, 256) == 103_388_545_024
    assert planner["x"].flops_per_forward(1, 256, 256) == 128_973_799_424
    assert planner["eps"].flops_per_forward(8, 512, 512) == 8 * 4 * 103_388_545_024


def test_workspace_scales_with_batch_and_reuses_buffers(planner):
    e = planner["eps"]
    w1, w8 = e.workspace_bytes(1, 256, 256), e.workspace_bytes(8, 256, 256)
    assert 6 * w1 < w8 < 9 * w1
    assert w1 < 100e6           # liveness-based reuse keeps one image-step under the 126 MB L2
    assert e.launches_per_step(8, 256, 256) == e.launches_per_forward(8, 256, 256) + 1


def test_planner_cannot_compute_and_rejects_bad_shapes(planner):
    from cdc_compression_b200 import EngineError
    e = planner["eps"]
    with pytest.raises(EngineError):
        e.workspace_bytes(1, 48, 64)
    with pytest.raises(EngineError):
        e.forward(torch.zeros(1, 3, 32, 32), torch.zeros(1), [])


def test_weight_validation(lib_built):
    from cdc_compression_b200 import DenoiserEngine, EngineError
    sd = O.seeded_unet_state_dict("eps", 0)
    e = DenoiserEngine("eps", 64, (1, 2, 3, 4, 5, 6), (1, 2, 3, 4), 3, 3, None)
    missing = dict(sd)
    del missing["downs.2.0.res_conv.weight"]
    with pytest.raises(EngineError, match="missing weight 'downs.2.0.res_conv.weight'"):
        e.load_weights(missing)
    bad = dict(sd)
    bad["ups.0.3.conv.weight"] = torch.zeros(320, 320, 3, 3)
    with pytest.raises(EngineError, match="ups.0.3.conv.weight"):
        e.load_weights(bad)
    with pytest.raises(EngineError, match="dim=32 unsupported"):
        DenoiserEngine("eps", 32, (1, 2), (1,), 3, 3, None)
    with pytest.raises(EngineError, match="CUDA"):
        DenoiserEngine("eps", 64, (1, 2), (1,), 3, 3, torch.device("cpu"))


@pytest.mark.parametrize("variant", ["eps", "x"])
def test_dropin_state_dict_is_the_reference_abi(variant):
    from conftest import build_dropin
    d = build_dropin(variant)
    sd = d.state_dict()
    shapes = O.unet_param_shapes(variant)
    mine = {k[len("denoise_fn."):]: tuple(v.shape) for k, v in sd.items() if k.startswith("denoise_fn.")}
    assert mine == shapes
    assert len(sd) == (472 if variant == "eps" else 473)
    for k in ("train_betas", "train_alphas_cumprod", "train_sqrt_recipm1_alphas_cumprod"):
        assert k in sd
    assert ("train_snr" in sd) == (variant == "x")
    T, sched = (20000, "linear") if variant == "eps" else (8193, "cosine")
    assert torch.equal(sd["train_alphas_cumprod"], O.train_alphas_cumprod(sched, T))
    # loss_fn_vgg.* (training-only LPIPS weights) are tolerated on load when lpips is absent
    extra = dict(sd)
    extra["loss_fn_vgg.lin0.model.1.weight"] = torch.zeros(1, 64, 1, 1)
    d.load_state_dict(extra)


@pytest.mark.parametrize("variant", ["eps", "x"])
def test_dropin_sample_schedule_matches_oracle(variant):
    from conftest import build_dropin
    d = build_dropin(variant, with_context_fn=False)
    T, sched = (20000, "linear") if variant == "eps" else (8193, "cosine")
    for S in (1, 9, 65):
        d.set_sample_schedule(S, torch.device("cpu"))
        sch = O.make_sample_schedule(O.train_alphas_cumprod(sched, T), S, variant)
        for k in ("alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "sigma"):
            assert torch.equal(torch.nan_to_num(getattr(d, k)), torch.nan_to_num(getattr(sch, k))), (S, k)
        coefs = d._coef_table(0.0)
        assert coefs.shape == (S, 8)
        for i in (0, S - 1):
            assert coefs[i, 5].item() == pytest.approx(O.unet_time(sch, i, variant, 1).item(), abs=0)


def test_no_cpu_path_in_dropin():
    from conftest import build_dropin
    from cdc_compression_b200 import EngineError
    d = build_dropin("eps", with_context_fn=False)
    x = torch.zeros(1, 3, 32, 32)
    with pytest.raises(EngineError, match="CUDA"):
        d.denoise_fn(x, torch.zeros(1, 1), O.seeded_context("eps", 1, 32, 32))


def test_ema_shim_is_a_checkpoint_container():
    sys.path.insert(0, os.path.join(ROOT, "cdc_compression_b200", "xparam"))
    try:
        sys.modules.pop("ema_pytorch", None)
        from ema_pytorch import EMA
    finally:
        sys.path.pop(0)
    m = torch.nn.Linear(2, 2)
    ema = EMA(m, beta=0.999, update_every=10, power=0.75, update_after_step=100)
    keys = set(ema.state_dict())
    assert {"ema_model.weight", "online_model.weight", "initted", "step"} <= keys
    sd = {k: torch.ones_like(v) for k, v in ema.state_dict().items()}
    ema.load_state_dict(sd)
    assert torch.equal(ema.ema_model.weight, torch.ones(2, 2))


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from cdc_compression_b200.parallel import shard_range, sharded_decode
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
torch.manual_seed(0)                       # every rank draws the SAME full-batch init before the split
images = torch.randn(5, 3, 8, 8); init = torch.randn(5, 3, 8, 8)
def decode(img, init=None, scale=2.0):     # stands in for diffusion.compress (per-image, batch-independent)
    return img * scale + init, img.flatten(1).sum(1)
out, bpp = sharded_decode(decode, images, init=init, scale=3.0)
ref, rbpp = decode(images, init=init, scale=3.0)
assert torch.equal(out, ref) and torch.equal(bpp, rbpp), "sharded result differs"
lo, hi = shard_range(5, dist.get_rank(), dist.get_world_size())
assert (hi - lo) in (2, 3)
dist.destroy_process_group()
print("ok")
'''


def test_batch_sharding_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, out


def test_shard_range_covers_batch():
    from cdc_compression_b200.parallel import shard_range
    for n in (1, 5, 8, 64):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_deepcopy_resets_engine_and_hook_marks_the_copy():
    """ADVICE r1 (CPU half; the GPU half is tests/test_gpu_parity_r2.py): copy.deepcopy of a drop-in model (what
    ema_pytorch.EMA does) must not alias the engine handle, and load_state_dict on the copy dirties the copy."""
    import copy
    from conftest import build_dropin
    d = build_dropin("eps", with_context_fn=False)
    d.denoise_fn._engine = object()          # stands in for a live DenoiserEngine
    d.denoise_fn._engine_dirty = False
    c = copy.deepcopy(d)
    assert c.denoise_fn._engine is None and c.denoise_fn._engine_dirty
    c.denoise_fn._engine_dirty = False
    c.load_state_dict(d.state_dict())        # through the parent module, as the demo scripts do
    assert c.denoise_fn._engine_dirty and not d.denoise_fn._engine_dirty
    d.denoise_fn._engine = None


def test_sharded_decode_rejects_shard_dependent_settings(monkeypatch):
    import functools
    import torch.distributed as dist
    from cdc_compression_b200 import parallel

    class FakeDiffusion:
        clip_noise = "half"

        def compress(self, images, init=None, **kw):
            return images, images.flatten(1).sum(1)

    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "get_rank", lambda group=None: 0)
    monkeypatch.setattr(dist, "get_world_size", lambda group=None: 2)
    imgs = torch.zeros(4, 3, 8, 8)
    with pytest.raises(NotImplementedError, match="half"):
        parallel.sharded_decode(functools.partial(FakeDiffusion().compress), imgs)
    ok = FakeDiffusion()
    ok.clip_noise = "none"
    with pytest.raises(NotImplementedError, match="eta"):
        parallel.sharded_decode(ok.compress, imgs, eta=0.5)
