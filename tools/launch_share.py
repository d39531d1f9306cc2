#!/usr/bin/env python
"""Share table of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`): time per kernel name,
share of the total, launches.  usage: python tools/launch_share.py x.csv [divide_by]  (divide_by = forwards / steps in the list)"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1], errors="ignore") if l.startswith('"')))
hdr = rows[0]
ci = {n: i for i, n in enumerate(hdr)}
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    if len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("spa3d::", "").replace("void ", "")
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print(f"total {total / div:.3f} ms per unit ({total:.2f} ms in the list, {sum(cnt.values())} launches)")
print("| kernel | share | per unit | launches |")
print("|---|---|---|---|")
for name, ms in tot.most_common(18):
    print(f"| `{name[:90]}` | {100 * ms / total:4.1f} % | {ms / div:6.3f} ms | {cnt[name] / div:.0f} |")
