#!/usr/bin/env python
"""bench.py — denoising throughput of the CDC decoder hot path on B200 (and the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA engine
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): epsilonparam variant, DDIM sampling loop of a 500-entry schedule,
batch = 8 images of 256x256 per GPU, eta = 0, clip_noise "none", seeded random-init weights (no checkpoint
is reachable offline), synthetic inputs.  A "step" is one DDIM step of the whole batch = one U-Net forward
(143 kernel launches) + the fused eps->x0->x_{t-1} update.  Metric: image-steps/s = images * steps / time
(pixel-steps/s is the same number x 65536).  Multi-GPU: each rank decodes its own 8 images (batch split,
no data-path collective) -> weak scaling; the time is the max over ranks of CUDA-event time.

One JSON line is printed by rank 0.  Extra objects: roofline (dominant kernel, measured live with
CUDA events), cpu_baseline (oracle port on the host cores, bounded sample), e2e (host buffers in/out
through diffusion.compress()), clocks, gpu_launches, blocks (per-op-family time share).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VARIANT = "eps"
BATCH, HEIGHT, WIDTH, SCHEDULE = 8, 256, 256, 500
PIX = HEIGHT * WIDTH


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        smax = max((int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=0)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(device):
    from conftest import build_dropin
    from oracle import cdc_oracle as O
    torch.manual_seed(0)
    d = build_dropin(VARIANT)
    sd = d.state_dict()
    for k, v in O.seeded_unet_state_dict(VARIANT, 0, gain=0.5).items():
        sd["denoise_fn." + k] = v
    d.load_state_dict(sd)
    return d.to(device)


def synthetic_batch(seed):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(BATCH, 3, HEIGHT, WIDTH, generator=g) * 2 - 1
    init = torch.randn(BATCH, 3, HEIGHT, WIDTH, generator=g) * 0.8
    return images, init


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_steps(n_steps, b, hw, warm=1):
    """Time `n_steps` DDIM steps of a b x 3 x hw x hw batch with the oracle (fp32, all host threads)."""
    from oracle import cdc_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.seeded_unet_state_dict(VARIANT, 0, gain=0.5)
    ctx = O.seeded_context(VARIANT, b, hw, hw)
    sch = O.make_sample_schedule(O.train_alphas_cumprod("linear", 20000), SCHEDULE, VARIANT)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(b, 3, hw, hw, generator=g) * 0.8
    with torch.no_grad():
        idx = list(range(SCHEDULE - 1, SCHEDULE - 1 - warm, -1))
        x = O.sample_loop(sd, sch, VARIANT, ctx, x, steps=idx)
        t0 = time.perf_counter()
        idx = list(range(SCHEDULE - 1 - warm, SCHEDULE - 1 - warm - n_steps, -1))
        O.sample_loop(sd, sch, VARIANT, ctx, x, steps=idx)
        dt = time.perf_counter() - t0
    return dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = 150.0
    t_probe = cpu_steps(1, 1, 64, warm=1)                    # seconds per 64x64 image-step
    ladder = [(8, 256), (4, 256), (2, 256), (1, 256), (1, 128), (1, 64)]
    b, hw = ladder[-1]
    for cand in ladder:
        est = t_probe * cand[0] * (cand[1] / 64) ** 2 * (args.steps + args.warmup)
        if est <= budget:
            b, hw = cand
            break
    dt = cpu_steps(args.steps, b, hw, warm=max(1, args.warmup))
    images_equiv = b * hw * hw / PIX
    value = images_equiv * args.steps / dt
    cores = torch.get_num_threads()
    sample = f"{args.steps} DDIM steps of a {b}x3x{hw}x{hw} batch (={images_equiv:g} images of 256x256 per step), fp32"
    line = {
        "impl": "reference", "metric": "denoising_image_steps_per_s", "value": value, "unit": "image-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "mpix_per_s": value * PIX / 1e6,
        "cpu_baseline": {"value": value, "unit": "image-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "image-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(workspace_mb=None):
    ws = f"{workspace_mb:.0f} MB workspace" if workspace_mb else "the engine workspace (~460 MB)"
    return {"workload": f"epsilonparam DDIM decode, {SCHEDULE}-entry schedule, batch={BATCH} {HEIGHT}x{WIDTH} per GPU "
                        "(BASELINE.json configs[1]), eta=0, clip_noise=none, seeded random-init weights",
            "variant": VARIANT, "batch_per_gpu": BATCH, "height": HEIGHT, "width": WIDTH, "schedule_steps": SCHEDULE,
            "l2_policy": f"working set per step ({ws} + ~100 MB fp16 weights) exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CDC engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = load_peaks()
    K, Wm = args.steps, args.warmup

    model = build_model(device)
    images, init = synthetic_batch(100 + rank)
    images_d, init_d = images.to(device), init.to(device)
    with torch.no_grad():
        ctx = model.context_fn(images_d, None)["output"]
    model.set_sample_schedule(SCHEDULE, device)
    x = init_d.clone().contiguous()
    eng = model._bind(x, ctx, 0.0)
    eng.set_context(ctx, BATCH, HEIGHT, WIDTH)
    flops_step = eng.flops_per_forward(BATCH, HEIGHT, WIDTH)
    launches_step = eng.launches_per_step(BATCH, HEIGHT, WIDTH)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then exactly K steps between CUDA events ----
    cursor = [SCHEDULE - 1]

    def run_steps(n):
        # n consecutive DDIM steps of the 500-entry schedule; a decode that reaches i = 0 restarts at i = S-1
        while n > 0:
            seg = min(n, cursor[0] + 1)
            eng.sample_loop(x, cursor[0], cursor[0] - seg + 1, "noise", "none")
            cursor[0] -= seg
            if cursor[0] < 0:
                cursor[0] = SCHEDULE - 1
            n -= seg

    run_steps(Wm)                                                  # warm-up (first step eager, graph captured)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    run_steps(K)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    value = world * BATCH * K / (ms_max / 1e3)
    finite = bool(torch.isfinite(x).all())

    # ---- end to end: host (pinned) images + init -> compress() -> host result, K-step decode ----
    images_p, init_p = images.pin_memory(), init.pin_memory()
    out_host = torch.empty_like(images_p).pin_memory()

    def e2e_once(steps):
        a = images_p.to(device, non_blocking=True)
        b = init_p.to(device, non_blocking=True)
        out, bpp = model.compress(a, sample_steps=steps, sample_mode="ddim", bpp_return_mean=False, init=b)
        out_host.copy_(out, non_blocking=True)
        return bpp

    Ke = min(K, SCHEDULE)                                          # one decode of Ke steps, host buffers both ends
    e2e_once(max(2, min(Wm, 8)))                                   # warm
    barrier()
    ev0.record()
    e2e_once(Ke)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * Ke / (t.item() / 1e3)
    h2d = (images_p.numel() + init_p.numel()) * 4 / Ke
    d2h = out_host.numel() * 4 / Ke

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-op timing (each op alone, CUDA events) -> dominant kernel + per-block table ----
    torch.cuda.synchronize()
    model.set_sample_schedule(SCHEDULE, device)
    xs = init_d.clone().contiguous()
    eng = model._bind(xs, ctx, 0.0)
    eng.set_context(ctx, BATCH, HEIGHT, WIDTH)
    eng.ddim_step(xs, SCHEDULE - 1, None, "noise", "none")
    prof = eng.profile_ops(iters=5)
    if args.ops_out:
        with open(args.ops_out, "w") as f:
            json.dump([{"op": n, "ms": m, "gflop": fl / 1e9, "tflops": (fl / (m * 1e-3) / 1e12) if m > 0 else 0.0}
                       for n, m, fl in prof], f, indent=0)
    tot_ms = sum(p[1] for p in prof)
    fam = {}
    for name, pms, fl in prof:
        base = name.replace("#partials", "")
        key = ("attention" if base.rsplit(".", 1)[-1] in ("ctx", "combine", "T", "M", "finish", "out") else
               "resample" if (base.endswith(".down") or base.endswith(".up")) else
               "res_conv" if base.endswith("res_conv") else
               "block_conv" if ("block1" in base or "block2" in base) else base)
        f = fam.setdefault(key, [0.0, 0.0])
        f[0] += pms
        f[1] += fl
    # dominant kernel = igemm_tc_kernel (the tcgen05/TMA implicit-GEMM convolution): every launch of it in one step
    # (Block convs, res_conv, Down/Upsample; sliced layers include their ln_rows_kernel second half).
    conv_ms = sum(v[0] for k, v in fam.items() if k in ("block_conv", "res_conv", "resample"))
    conv_fl = sum(v[1] for k, v in fam.items() if k in ("block_conv", "res_conv", "resample"))
    conv_tf = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    top = max((p for p in prof if p[2] > 0), key=lambda p: p[2] / max(p[1], 1e-9))
    traffic = None
    try:   # DRAM bytes per launch of the named instance, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")) as f:
            traffic = json.load(f).get("igemm_tc_kernel")
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": conv_tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                "frac": conv_tf / peaks["tf_burst"], "traffic": traffic, "kernel": "igemm_tc_kernel",
                "launches_per_step": sum(1 for p in prof if p[2] > 0 and not p[0].endswith(".ctx") and p[0] != "final_conv"),
                "algorithmic_flops_per_step": conv_fl, "kernel_ms_per_step": conv_ms,
                "kernel_share_of_step": conv_ms / tot_ms if tot_ms else None,
                "best_launch": {"op": top[0], "tflops": top[2] / (top[1] * 1e-3) / 1e12, "ms": top[1]},
                "peak_source": peaks["source"] + " burst (launches timed alone, CUDA events)"}
    step_tf = flops_step * K / (ms_max * 1e-3) / 1e12
    line = {
        "metric": "denoising_image_steps_per_s", "value": value, "unit": "image-steps/s", "n_gpus": world,
        "steps": K, "warmup": Wm, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(eng.workspace_bytes(BATCH, HEIGHT, WIDTH) / 1e6),
        "mpix_per_s": value * PIX / 1e6, "batch_steps_per_s": K / (ms_max / 1e3),
        "roofline": roofline,
        "step_roofline": {"bound": "tensor", "achieved": step_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                          "frac": step_tf / peaks["tf_sustained"], "flops_per_step": flops_step,
                          "peak_source": peaks["source"] + " sustained (whole step)"},
        "blocks": {k: {"ms": v[0], "share": v[0] / tot_ms, "tflops": (v[1] / (v[0] * 1e-3) / 1e12) if v[0] else 0.0}
                   for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])},
        "e2e": {"value": e2e_value, "unit": "image-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": f"pinned host images+init -> H2D -> context_fn -> {K}-step DDIM decode -> D2H, via "
                        "GaussianDiffusion.compress()"},
        "gpu_launches": launches_step * K * world, "clocks": clocks, "finite": finite,
    }
    if not args.no_cpu_baseline and world == 1:
        n_cpu = 3
        dt = cpu_steps(n_cpu, BATCH, 256, warm=1)
        line["cpu_baseline"] = {"value": BATCH * n_cpu / dt, "unit": "image-steps/s", "cores": torch.get_num_threads(),
                                "kind": "port",
                                "sample": f"{n_cpu} DDIM steps (+1 warm-up) of the same {BATCH}x3x256x256 batch with the "
                                          f"oracle port, fp32, {torch.get_num_threads()} threads, {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ops-out", default=None, help="write the per-op timing table (JSON) here")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
