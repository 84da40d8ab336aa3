#!/usr/bin/env python3
"""Compare the engine's merge list with the oracle's for several batch caps (debug helper)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import llmtokenizer_b200 as L
import oracle_api
from llmtokenizer_b200 import _lib
lib = _lib.load_corpus()
kind, size, seed, cap = 0, 12_000_000, 71, 2500
data = np.zeros(size, dtype=np.uint8)
lib.gen_corpus_fill(kind, data.ctypes.data, size, seed, 50000)
oracle = oracle_api.load()
rc, om, ot, ost = oracle.train(data, cap, oracle_api.FAST_CF)
for bm in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4, 8]:
    ctx = L.Context(0)
    ctx.set_option("batch_max", bm)
    ctx.upload(data)
    st = ctx.train(cap)
    m, t = ctx.download()
    ctx.close()
    bad = next((i for i in range(min(len(m), len(om))) if tuple(m[i]) != tuple(om[i])), None)
    print("batch_max", bm, "passes", st["replace_passes"], "first mismatch", bad, flush=True)
    if bad is not None:
        lo, hi = max(0, bad - 3), bad + 6
        print("  engine", [tuple(int(v) for v in x) for x in m[lo:hi]])
        print("  oracle", [tuple(int(v) for v in x) for x in om[lo:hi]])
        es = set(map(tuple, m.tolist())); os_ = set(map(tuple, om.tolist()))
        print("  same set of merges:", es == os_, "ids equal:", np.array_equal(t, ot))
