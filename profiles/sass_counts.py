"""Per-kernel SASS evidence for the built library (run in the build container, no GPU needed):

    python profiles/sass_counts.py cdc_compression_b200/libcdc_b200.so > profiles/sass_r02.txt

Counts, per kernel of libcdc_b200.so, the mnemonics that prove (or disprove) a Blackwell-native path
(B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store,
UTMACCTL.PF = tensor-map prefetch, SYNCS = mbarrier, HMMA = legacy mma.sync, LDGSTS = cp.async, REDG/ATOMG = global atomics.
"""
import re
import subprocess
import sys
from collections import OrderedDict

lib = sys.argv[1]
names = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMACCTL", "SYNCS", "HMMA", "LDGSTS", "ATOMG", "REDG", "FFMA2", "MUFU"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern = OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = dict.fromkeys(names, 0)
        kern[cur]["instrs"] = 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    kern[cur]["instrs"] += 1
    for n in names:
        if op.startswith(n):
            kern[cur][n] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib}: per-kernel SASS mnemonic counts (static instruction counts, not executions)")
print("| kernel | instrs | " + " | ".join(names) + " |")
print("|---|---|" + "---|" * len(names))
for (k, c), d in zip(kern.items(), demangled):
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("cdc::", "")
    short = re.sub(r"\(int\)|\(bool\)", "", short)
    print(f"| {short} | {c['instrs']} | " + " | ".join(str(c[n]) for n in names) + " |")
