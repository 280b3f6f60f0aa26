"""Warp-stall breakdown and hottest instructions from `ncu --set full --import-source on` captures (build container):

    python profiles/summarize_ncu_source.py gpurun_out/r2/prof_l0.ncu-rep gpurun_out/r2/prof_l1.ncu-rep > profiles/ncu_r02_source_level.md
"""
import csv
import subprocess
import sys


def sections(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows, secs, cur = list(csv.reader(raw.splitlines())), [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    # ncu prints every kernel twice (SASS and source views): keep the first of each launch
    seen, out = set(), []
    for s in secs:
        key = (s["name"], len(s["rows"]))
        if key in seen:
            continue
        seen.add(key)
        out.append(s)
    return out


print("# ncu source-level view of the convolution kernel (round 2, B = 8, 256 x 256, eps)\n")
print("`ncu --set full --clock-control none --import-source on -k regex:igemm_tc_kernel -s <n> -c <m> python "
      "tests/gpu_profile_forward.py --iters 1`; launches: downs.0.1.block1 (64->64, LayerNorm+ReLU+shift, row in registers), "
      "downs.0.1.block2 (+ identity residual through the second TMEM accumulator, hi/lo stores), downs.1.1.block1 (128->128).\n"
      "Samples cover all six warps of a CTA: the `BRA` rows with stall_long_sb are mbarrier polls (producer waiting for a free "
      "stage, MMA issuer waiting for data, epilogue waiting for the accumulator).\n")
for rep in sys.argv[1:]:
    for s in sections(rep):
        h = s["hdr"]
        ix = {n: i for i, n in enumerate(h)}
        stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        R = [r for r in s["rows"] if len(r) >= len(h)]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in R)
        agg = {n: sum(int(r[ix[n]] or 0) for r in R) for n in stall}
        short = s["name"].split("(cdc::")[0].replace("void cdc::", "").replace("(int)", "").replace("(bool)", "")
        print(f"## {short} — {tot} samples\n")
        print("| stall reason | samples | share |\n|---|---|---|")
        for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
            print(f"| {n} | {v} | {100 * v / max(tot, 1):.1f} % |")
        print("\n| samples | executed | instruction | top stall |\n|---|---|---|---|")
        for r in sorted(R, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
            st = max(((int(r[ix[n]] or 0), n) for n in stall))
            print(f"| {r[ix['# Samples']]} | {r[ix['Instructions Executed']]} | `{r[ix['Source']].strip()[:60]}` | {st[1]} {st[0]} |")
        print()
