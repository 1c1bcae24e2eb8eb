"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (development aid / profiles/)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
seq = []
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
    val = float(row["Metric Value"].replace(",", ""))
    us = val / 1000.0 if row["Metric Unit"].startswith("n") else val
    seq.append((row["ID"], name, us, row.get("Grid Size"), row.get("Block Size")))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {len(seq)} launches, {tot/1e3:.3f} ms total (ncu: serialised, cold cache -- compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:56]:56s} n={n:5d} total={t:10.1f} us avg={t/n:9.2f} us share={t/tot:.3f}")
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for s in seq[a:b]:
        print(s)
