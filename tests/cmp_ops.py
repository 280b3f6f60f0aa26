"""Compare two per-op timing tables (bench.py --ops-out): python tests/cmp_ops.py old.json new.json [min_us]"""
import json
import sys

a = {o["op"]: o for o in json.load(open(sys.argv[1]))}
b = json.load(open(sys.argv[2]))
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
ta = tb = 0.0
for o in b:
    n = o["op"]
    if n not in a:
        print(f"{n:36s}      new {o['ms']*1e3:7.1f}")
        tb += o["ms"]
        continue
    x, y = a[n]["ms"] * 1e3, o["ms"] * 1e3
    ta += x / 1e3
    tb += y / 1e3
    if abs(x - y) >= thr:
        print(f"{n:36s} {x:7.1f} -> {y:7.1f}  {y - x:+6.1f}")
print(f"sum {ta*1e3:.1f} -> {tb*1e3:.1f} us")
