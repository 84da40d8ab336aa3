#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: the hottest SASS instructions of the first profiled
launch with their stall reasons.  usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_hot.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data, seen = [], False
for r in rows:
    if r and r[0] == "Address":
        if seen:
            break
        seen = True
        continue
    if seen and len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {k: sum(int(r[ix[k]]) for r in data) for k in keys}
print("by reason:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:N]
for i in sorted(top):
    r = data[i]
    st = {k[6:]: int(r[ix[k]]) for k in keys if int(r[ix[k]])}
    print(f"{i:5d} {r[ix['Source']].strip()[:84]:84s} {r[ix['# Samples']]:>6s} {st}")
