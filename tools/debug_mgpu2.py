#!/usr/bin/env python3
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llmtokenizer_b200 as L
from llmtokenizer_b200 import _lib
P = int(sys.argv[1])
lib = _lib.load_corpus()
buf = np.zeros(3_000_000, dtype=np.uint8)
lib.gen_corpus_fill(0, buf.ctypes.data, buf.size, 9, 50000)
m = np.load("/tmp/m.npy") if os.path.exists("/tmp/m.npy") else None
if m is None:
    m, t, _ = L.train(buf, max_merges=300)
    np.save("/tmp/m.npy", m)
    sys.exit(0)
try:
    ids, st = L.encode(buf, m, n_gpus=P)
    print("ok", len(ids))
except Exception as e:
    print("FAILED", e)
