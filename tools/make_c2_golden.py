#!/usr/bin/env python3
"""Full-size parity fixture for config 2: run the CPU oracle (oracle/, pinned against the compiled reference)
on the exact corpus bench.py uses (zipf_ascii, 100 MB, seed 1234, 4,096 merges; ~4 minutes on one core) and
record the SHA-256 of the merge list and of the ids.  tests/test_gpu_parity.py and bench.py compare the
engine's result with these digests.

  python tools/make_c2_golden.py                                   # writes tests/golden/c2_full.json
  python tools/make_c2_golden.py zipf12m_10k 12000000 11 10000     # name, bytes, seed, merges [, kind]: other digests
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_api  # noqa: E402
from llmtokenizer_b200 import _lib  # noqa: E402

SIZE, SEED, VOCAB, MERGES = 100_000_000, 1234, 50000, 4096


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u4").tobytes()).hexdigest()


def main():
    global SIZE, SEED, MERGES
    name, kind = "c2_full", 0
    if len(sys.argv) >= 5:
        name, SIZE, SEED, MERGES = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        kind = int(sys.argv[5]) if len(sys.argv) > 5 else 0      # 0 = zipf_ascii, 1 = zipf_bytes (config 3)
    buf = np.zeros(SIZE, dtype=np.uint8)
    assert _lib.load_corpus().gen_corpus_fill(kind, buf.ctypes.data, SIZE, SEED, VOCAB if kind == 0 else 65536) == 0
    t0 = time.time()
    rc, merges, ids, _ = oracle_api.load().train(buf, MERGES, oracle_api.FAST_CF)
    assert rc == 0 and len(merges) == MERGES
    out = {"corpus": {"kind": "zipf_ascii" if kind == 0 else "zipf_bytes", "bytes": SIZE, "seed": SEED,
                      "words": VOCAB if kind == 0 else 65536}, "merges": MERGES,
           "n_ids": int(len(ids)), "merges_sha256": sha(merges), "ids_sha256": sha(ids),
           "made_by": "tools/make_c2_golden.py (oracle FAST_CF mode)", "oracle_seconds": round(time.time() - t0, 1)}
    with open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
