"""Aggregate an ncu gpu__time_duration launch list by kernel name.  usage: agg_launches.py launches.csv [skip_first_n]"""
import collections, csv, re, sys
with open(sys.argv[1]) as f:
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
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = data[skip:]
tot = sum(v for _, v in data)
short = lambda k: re.sub(r"\(.*", "", k).replace("void ", "").replace("<unnamed>::", "")[:90]
agg = collections.OrderedDict()
for k, v in data:
    a = agg.setdefault(short(k), [0, 0.0])
    a[0] += 1
    a[1] += v
print(f"{len(data)} launches, {tot / 1e3:.1f} us (ncu, serialised)")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v / 1e3:9.1f} us {100 * v / tot:5.1f}% x{n:4d} ({v / n / 1e3:8.1f} us each)  {k}")
