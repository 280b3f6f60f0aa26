"""Turn an `ncu --set full` report into the small tables committed under profiles/ (run in the build container):

    python profiles/summarize_ncu.py gpurun_out/prof_r01_tc.ncu-rep profiles/ncu_r01_igemm_tc.md [names...]

`names` optionally labels the captured launches in order (plan op names).  Also prints the DRAM traffic per launch.
"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
names = sys.argv[3:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def col(r, name, default=""):
    return r[ix[name]] if name in ix else default


cols = [
    ("kernel", "Kernel Name"), ("grid", "Grid Size"), ("us", "gpu__time_duration.sum"),
    ("regs", "launch__registers_per_thread"), ("dram rd MB", "dram__bytes_read.sum"),
    ("dram wr MB", "dram__bytes_write.sum"), ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
]
lines = ["| # | op | " + " | ".join(c[0] for c in cols) + " |", "|" + "---|" * (len(cols) + 2)]
for n, r in enumerate(rows[2:]):
    vals = []
    for label, key in cols:
        v = col(r, key)
        if label == "kernel":
            v = v.split("(")[0].replace("void ", "")
        else:
            try:
                v = f"{float(v.replace(',', '')):.1f}"
            except ValueError:
                pass
        vals.append(v)
    lines.append(f"| {n} | {names[n] if n < len(names) else ''} | " + " | ".join(vals) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
