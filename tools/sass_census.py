#!/usr/bin/env python
"""Per-kernel census of the Blackwell-specific SASS in lib3dspa_b200.so: UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTMALDG / UTMASTG (TMA loads / stores), HMMA (mma.sync - none expected), MUFU.  usage: python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "3dspa_code_b200", "lib3dspa_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU")
cur, cnt, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        cnt[cur]["instructions"] += 1
        for k in KEYS:
            if op == k or (k != "HMMA" and op.startswith(k)):
                cnt[cur][k] += 1
                total[k] += 1
names = subprocess.run(["cu++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); HMMA = mma.sync tensor-core instructions (legacy path)")
print("# totals: " + ", ".join(f"{k} {total[k]}" for k in KEYS) + f"; kernels {len(cnt)}")
print(f"{'kernel':110s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in KEYS))
for (mangled, c), name in zip(cnt.items(), names):
    i = name.rfind(">(")
    short = (name[: i + 1] if i >= 0 else name.split("(")[0]).replace("spa3d::", "")
    short = re.sub(r"\((int|bool)\)", "", short)
    print(f"{short[:110]:110s} {c['instructions']:7d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
