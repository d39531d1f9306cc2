#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall reasons, hottest SASS instructions, opcode mix.
usage: ncu -i rep.ncu-rep --page source --csv [--kernel-id ...] > src.csv; python tools/ncu_src_summary.py src.csv [first-kernel-only]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = starts[0]
end = starts[1] - 1 if len(starts) > 1 else len(rows)
hdr = rows[h]
data = [r for r in rows[h + 1 : end] if len(r) == len(hdr) and r[0].startswith("0x")]
ci = {n: i for i, n in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]]) for r in data)
print("kernel:", rows[h - 1][1][:120] if h > 0 else "?")
print("total samples", tot, "instructions", len(data))
reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ci[n]]) for r in data) for n in reasons}
for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {n:26s} {v:8d} {100 * v / max(tot, 1):5.1f}%")
print("hottest instructions:")
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    rs = {n: int(r[ci[n]]) for n in reasons if int(r[ci[n]]) > 0}
    main = sorted(rs.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {int(r[ci['# Samples']]):6d} {r[ci['Source']].strip()[:72]:72s} {main}")
op, opc = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
    o = m.group(2).split(".")[0] if m else "?"
    op[o] += int(r[ci["# Samples"]])
    opc[o] += int(r[ci["Instructions Executed"]])
print("samples by opcode:", op.most_common(14))
print("executed by opcode:", opc.most_common(14))
