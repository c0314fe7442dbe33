"""Per-kernel breakdown of ONE training step from an ncu launch list (gpu__time_duration.sum CSV of `bench.py`).
usage: step_breakdown.py launches.csv [--list]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = []
for x in r:
    if len(x) > iv:
        try:
            data.append((x[ik], float(x[iv].replace(",", ""))))
        except ValueError:
            pass
idx = [i for i, d in enumerate(data) if d[0].startswith("k_adam(") or d[0] == "k_adam"]
seg = data[idx[-3] + 1: idx[-2] + 1] if len(idx) >= 3 else data
tot = sum(v for _, v in seg)
print(f"step: {len(seg)} launches, {tot / 1e3:.1f} us (ncu, serialised, cold cache)")
short = lambda k: re.sub(r"\(.*", "", k).replace("void ", "").replace("<unnamed>::", "")[:80]
if "--list" in sys.argv:
    for k, v in seg:
        print(f"{v / 1e3:8.1f}  {short(k)}")
agg = collections.OrderedDict()
for k, v in seg:
    a = agg.setdefault(short(k), [0, 0.0])
    a[0] += 1
    a[1] += v
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / 1e3:9.1f} us {100 * v / tot:5.1f}% x{n:3d}  {k}")
