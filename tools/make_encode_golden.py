#!/usr/bin/env python3
"""Encode fixture for config 4's shape: one 125 MB chunk of its corpus (zipf_bytes, sampling seed 777 + k, config 3's word
list) encoded with config 3's committed 32,000-merge table by the CPU oracle (bo_encode: bpe.c:760-772 rank by rank, one
thread, about an hour).  Writes tests/golden/<name>.json (digest of the ids).

  python tools/make_encode_golden.py c4_chunk0_encode 125000000 777
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_api  # noqa: E402
from parity_cases import corpus, sha  # noqa: E402


def main():
    name, size, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    table = "tests/golden/full/c3_full_merges.npz"
    merges = np.load(os.path.join(ROOT, table))["merges"]
    data = corpus(1, size, seed)
    t0 = time.time()
    ids = oracle_api.load().encode(data, merges)
    out = {"corpus": {"kind": "zipf_bytes", "bytes": size, "seed": seed, "words": 65536},
           "table": table + " (config 3's 32,000 merges)", "ranks": int(len(merges)), "n_ids": int(len(ids)),
           "ids_sha256": sha(ids), "made_by": "tools/make_encode_golden.py (oracle bo_encode, 1 thread)",
           "oracle_seconds": round(time.time() - t0, 1)}
    with open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
