"""Train on a synthetic corpus with batched passes on and off and save both merge lists (debugging aid).
  python tools/dump_merges.py KIND BYTES SEED MERGES OUT.npz"""
import sys

import numpy as np

sys.path.insert(0, ".")
import llmtokenizer_b200 as L
from llmtokenizer_b200 import _lib

kind, size, seed, merges, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
buf = np.zeros(size, dtype=np.uint8)
assert _lib.load_corpus().gen_corpus_fill(kind, buf.ctypes.data, size, seed, 50000 if kind == 0 else 65536) == 0
res = {}
for bm in (8, 1):
    ctx = L.Context(0)
    ctx.set_option("batch_max", bm)
    ctx.upload(buf)
    st = ctx.train(merges)
    m, t = ctx.download()
    nb = ctx.decode(m, download=False)
    print(f"batch_max {bm}: merges {len(m)} ids {len(t)} passes {st['replace_passes']} batched {st['batch_merges']} ties {st['same_bucket_ties']} "
          f"edges {st['threshold_edges']} rehashes {st['table_rehashes']} decode bytes {nb} mismatches {ctx.decode_mismatches()}")
    ctx.close()
    res[f"m{bm}"] = m
    res[f"n{bm}"] = np.array([len(t)])
k = min(len(res["m8"]), len(res["m1"]))
bad = next((i for i in range(k) if tuple(res["m8"][i]) != tuple(res["m1"][i])), None)
print("first difference between batched and unbatched:", bad, None if bad is None else (res["m8"][bad], res["m1"][bad]))
np.savez_compressed(out, **res)
