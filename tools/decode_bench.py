"""Time the on-device decode of a trained stream (python tools/decode_bench.py [bytes] [merges])."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import llmtokenizer_b200 as L
from llmtokenizer_b200 import _lib

size = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lib = _lib.load_corpus()
buf = np.zeros(size, dtype=np.uint8)
assert lib.gen_corpus_fill(0, buf.ctypes.data, size, 1, 50000) == 0
ctx = L.Context(0)
ctx.upload(buf)
st = ctx.train(cap)
m, _ = ctx.download(tokens=False)
print("trained", st["n_merges"], "merges,", st["n_tokens"], "ids")
for it in range(5):
    t0 = time.perf_counter()
    nb = ctx.decode(m, download=False)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    bad = ctx.decode_mismatches()
    dc = time.perf_counter() - t1
    print(f"decode {nb} bytes in {dt*1e3:.2f} ms ({nb/dt/1e9:.1f} GB/s of output, wall incl. vocabulary flatten + sync); "
          f"compare {dc*1e3:.2f} ms, mismatches {bad}")
ctx.close()
