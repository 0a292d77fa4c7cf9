"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of GPU time)."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"launches {sum(v[0] for v in agg.values())}  total {tot:.1f} us (cold-cache, serialised: compare SHARES)")
tc = sum(t for k, (n, t) in agg.items() if k.startswith("tc::"))
print(f"tcgen05 contraction kernels (tc::*): {100 * tc / tot:.1f}% of GPU time")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:5d} {t:10.1f} us {100 * t / tot:5.1f}%  {k[:120]}")
