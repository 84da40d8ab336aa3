#!/usr/bin/env python3
"""Bisect helper: sharded encode / train against the single-GPU result (engine options come from the environment)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llmtokenizer_b200 as L
from llmtokenizer_b200 import _lib

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lib = _lib.load_corpus()
def corpus(kind, size, seed):
    buf = np.zeros(size, dtype=np.uint8)
    lib.gen_corpus_fill(kind, buf.ctypes.data, size, seed, 50000 if kind == 0 else 65536)
    return buf
data = corpus(0, 3_000_000, 9)
m, t, _ = L.train(data, max_merges=300)
print("trained", len(m), len(t), flush=True)
for what, d in (("same", data), ("other", corpus(0, 2_000_001, 10))):
    try:
        ref, _ = L.encode(d, m, n_gpus=1)
        ids, st = L.encode(d, m, n_gpus=P)
        print(what, "P", P, "equal:", np.array_equal(ids, ref), len(ids), len(ref), flush=True)
    except Exception as e:
        print(what, "FAILED", e, flush=True)
