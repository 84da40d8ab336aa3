#!/usr/bin/env python3
"""How much would warp-level aggregation (__match_any_sync + one RED per distinct address) save on the global
pair-count deltas of replace_stream_kernel<false>?  CPU-side count on the real token streams (the oracle's),
no GPU needed.

The kernel's scanners work on 128-token warp-iterations; every replacement that starts in one emits up to four
reductions ((x,a) gone, (x',z) new, (b,y) gone, (z,y) new) into the L2-resident delta vectors.  Aggregation pays
only where two reductions of one warp-iteration hit the SAME counter.  For the merge that follows `after` merges
this tool reports: replacements per warp-iteration, the share of warp-iterations that have any, and
duplicates / reductions (what aggregation could remove) against the 4 MATCH + 4 REDUX warp instructions it would
add to every warp-iteration that has a replacement.

  python tools/warp_dup_stats.py c2 192 1000 4000        # workload (bench.WORKLOADS, at most 200 MB is used), merges done
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import oracle_api  # noqa: E402

SENT = 0xFFFFFFFF


def stats(ids, a, b, z):
    n = ids.size
    m = np.flatnonzero((ids[:-1] == a) & (ids[1:] == b))        # a != b: every occurrence is a replacement
    if a == b or m.size == 0:
        return None
    pad = np.concatenate([[SENT, SENT], ids, [SENT, SENT, SENT]]).astype(np.int64)
    p = m + 2
    xl, xll, yr, yrr = pad[p - 1], pad[p - 2], pad[p + 2], pad[p + 3]
    pm = (xll == a) & (xl == b)
    nm = (yr == a) & (yrr == b)
    win = m // 128
    addr, w = [], []
    ok = xl != SENT
    addr += [xl[ok] * 4 + 0, np.where(pm, z, xl)[ok] * 4 + 2]
    w += [win[ok], win[ok]]
    ok = (yr != SENT) & ~nm
    addr += [yr[ok] * 4 + 1, yr[ok] * 4 + 3]
    w += [win[ok], win[ok]]
    addr = np.concatenate(addr)
    w = np.concatenate(w)
    key = w.astype(np.int64) * (4 * (z + 1)) + addr
    distinct = np.unique(key).size
    windows = (n + 127) // 128
    busy = np.unique(win).size
    return {"tokens": int(n), "replacements": int(m.size), "per_warp_iteration": m.size / windows,
            "busy_share": busy / windows, "reductions": int(key.size), "duplicates": int(key.size - distinct),
            "dup_share": (key.size - distinct) / key.size, "extra_warp_instr": int(8 * busy)}


def main():
    w = dict(bench.WORKLOADS[sys.argv[1]])
    w["size"] = min(w["size"], 200_000_000)
    afters = [int(x) for x in sys.argv[2:]] or [192, 1000, 4000]
    data = np.empty(w["size"], dtype=np.uint8)
    bench.fill_corpus(w, data, 0, w["size"])
    orc = oracle_api.load()
    orc.configure(workers=os.cpu_count() or 1)
    rc, merges, _, _ = orc.train(data, max(afters) + 1, oracle_api.FAST_CF)
    assert rc == 0
    for k in afters:
        rc, _, ids, _ = orc.train(data, k, oracle_api.FAST_CF)   # the stream after k merges
        assert rc == 0
        a, b = (int(v) for v in merges[k])
        s = stats(ids, a, b, 256 + k)
        print(f"{sys.argv[1]} ({w['size'] / 1e6:.0f} MB) merge #{k} ({a},{b})->{256 + k}:", s, flush=True)


if __name__ == "__main__":
    main()
