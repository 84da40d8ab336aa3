#!/usr/bin/env python3
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`):
launches, total time and share per kernel.  usage: python tools/launch_summary.py X.csv[.gz] [title]"""
import csv
import gzip
import re
import sys
from collections import defaultdict

path = sys.argv[1]
fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = [r for r in csv.reader(l for l in fh if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt, fall = defaultdict(float), defaultdict(int), defaultdict(int)
for r in rows[1:]:
    if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("bpe::", "").strip()
    tot[name] += v
    cnt[name] += 1
    fall[name] += v < 6.0
allt = sum(tot.values())
if len(sys.argv) > 2:
    print("#", sys.argv[2])
print(f"# {sum(cnt.values())} launches, {allt / 1e3:.2f} ms of kernel time (per-launch times under ncu are cold-cache and serialised: "
      f"compare SHARES; launches shorter than 6 us are speculative steps that fall through)\n")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:46s} n={cnt[k]:6d} total={tot[k] / 1e3:9.2f} ms share={100 * tot[k] / allt:5.1f}% avg={tot[k] / cnt[k]:8.1f} us  "
          f"fall-through={fall[k]}")
