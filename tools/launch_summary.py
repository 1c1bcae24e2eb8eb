#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total / mean device time, share.
    python tools/launch_summary.py gpurun_out/r2_launches_decode.csv > profiles/r2_launches_decode.summary.txt
(per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's `stages`, not absolutes)"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
col = {h: i for i, h in enumerate(rows[hdr])}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[hdr + 1:]:
    if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("wb::<unnamed>::", "")
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values()) or 1.0
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {total / 1e3:.3f} ms of device time (serialised, cold cache)")
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'mean us':>9s} {'share':>7s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {tot[k]:12.1f} {tot[k] / cnt[k]:9.2f} {100 * tot[k] / total:6.1f}%")
