"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tests/launch_summary.py file.csv [skip_first_n]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
rows = rows[skip:]
agg = collections.OrderedDict()
for r in rows:
    m = re.search(r"(pe_\w+)", r["Kernel Name"])
    k = m.group(1) if m else re.sub(r"<.*", "", r["Kernel Name"])[:36]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms (cold-cache, serialised)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{k:34s} {c:4d} {v / 1e6:9.3f} ms {100 * v / tot:5.1f}%")
