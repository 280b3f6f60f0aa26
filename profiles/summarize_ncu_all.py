"""Per-launch table of ONE U-Net forward from an ncu capture of every launch (run in the build container):

    python profiles/summarize_ncu_all.py gpurun_out/r2f/prof_all.ncu-rep gpurun_out/r2f/op_names.json \
        profiles/ncu_r02_all.md profiles/ncu_traffic_r02.json

The capture: `ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy
--clock-control none -s 144 -c 144 python tests/gpu_profile_forward.py --iters 2 --names-out ...` (B = 8, 256 x 256, eps;
the second of two forwards; cold-cache and serialised — compare shares, not absolutes).  Launches are labelled with the
plan's op names (aligned on the first time_mlp_kernel).  DRAM bytes per launch = dram__bytes.sum.per_second x duration.
"""
import csv
import json
import subprocess
import sys

rep, names_json, out_md, out_json = sys.argv[1:5]
names = json.load(open(names_json))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, key):
    try:
        return float(r[ix[key]].replace(",", ""))
    except (KeyError, ValueError):
        return None


def us(r):
    v, u = num(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]]
    return v / 1000 if u.startswith("n") else (v * 1000 if u.startswith("m") else v)


def dram_mb(r):
    v, u = num(r, "dram__bytes.sum.per_second"), units[ix["dram__bytes.sum.per_second"]]
    scale = {"Tbyte/s": 1e12, "Gbyte/s": 1e9, "Mbyte/s": 1e6, "Kbyte/s": 1e3, "byte/s": 1.0}[u]
    return v * scale * us(r) * 1e-6 / 1e6


kern = lambda r: r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("cdc::", "")
first = next(i for i, r in enumerate(data) if kern(r).startswith("time_mlp_kernel"))
cols = [("us", None), ("dram MB", None), ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread")]
lines = ["| # | op | kernel | grid | " + " | ".join(c[0] for c in cols) + " |", "|" + "---|" * (len(cols) + 4)]
per_op, total = {}, 0.0
for i, r in enumerate(data[first:]):
    if i >= len(names):
        break
    vals = []
    for label, key in cols:
        v = us(r) if label == "us" else dram_mb(r) if label == "dram MB" else num(r, key)
        vals.append("" if v is None else f"{v:.1f}")
    lines.append(f"| {i} | {names[i]} | {kern(r)} | {r[ix['Grid Size']]} | " + " | ".join(vals) + " |")
    base = names[i].replace("#partials", "")
    e = per_op.setdefault(base, {"us": 0.0, "dram_mb": 0.0})
    e["us"] += us(r)
    e["dram_mb"] += dram_mb(r)
    total += us(r)
lines.append("")
lines.append(f"Sum of the {min(len(names), len(data) - first)} launch durations: {total:.0f} us (serialised, cold caches; the step graph "
             "overlaps the prologues with programmatic dependent launch and runs warm).")
open(out_md, "w").write("\n".join(lines) + "\n")

groups = {
    "final_conv": ["final_conv"], "ups.4.3.up": ["ups.4.3.up"], "downs.0.3.down": ["downs.0.3.down"],
    "downs.0.2 attention": ["downs.0.2.ctx", "downs.0.2.combine", "downs.0.2.T", "downs.0.2.M", "downs.0.2.out"],
    "ups.4.2 attention": ["ups.4.2.ctx", "ups.4.2.combine", "ups.4.2.T", "ups.4.2.M", "ups.4.2.out"],
    "pack_input": ["pack_input"],
}
mem = {g: sum(per_op.get(o, {"dram_mb": 0})["dram_mb"] for o in ops) * 1e6 for g, ops in groups.items()}
top = per_op.get("downs.0.1.block2", {"us": 0, "dram_mb": 0})
json.dump({"memory_bound": mem,
           "igemm_tc_kernel": {"op": "downs.0.1.block2 (64->64 3x3 conv + LN + ReLU + identity residual through the MMA, 8x256x256)",
                               "dram_bytes": top["dram_mb"] * 1e6, "us_under_ncu": top["us"],
                               "algorithmic_bytes_hi_lo": 335.5e6, "algorithmic_bytes_single_fp16": 201.3e6,
                               "source": out_md + " (ncu, this round's kernels; bytes = dram__bytes.sum.per_second x duration)"},
           "per_op_dram_mb": {k: round(v["dram_mb"], 2) for k, v in per_op.items()}}, open(out_json, "w"), indent=1)
print("\n".join(lines[:8]))
print("...", len(lines), "lines; total us", round(total))
