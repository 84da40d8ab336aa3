#!/usr/bin/env python3
"""Short, profiler-friendly run of the hot path: one training (and optionally one encoding) of a
synthetic corpus through the C ABI.  Used under `ncu` on the GPU box (profiles/README.md lists the
exact command lines); prints the engine's own statistics as one JSON line.

  python tools/profile_run.py [--kind 0|1] [--size BYTES] [--seed S] [--merges M] [--repeat R] [--encode]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", type=int, default=0, help="0 = zipf_ascii (c2), 1 = zipf_bytes (c3)")
    ap.add_argument("--size", type=int, default=100_000_000)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--merges", type=int, default=128)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--encode", action="store_true")
    ap.add_argument("--profile-replace", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (repeatable)")
    args = ap.parse_args()

    import llmtokenizer_b200 as L
    from llmtokenizer_b200 import _lib
    corpus = _lib.load_corpus()
    if args.kind == 2:  # uniform random bytes 1..255: every pair is rare, the passes are pure streaming
        data = np.random.default_rng(args.seed).integers(1, 256, args.size, dtype=np.uint8)
    else:
        data = np.zeros(args.size, dtype=np.uint8)
        assert corpus.gen_corpus_fill(args.kind, data.ctypes.data, data.size, args.seed,
                                      50000 if args.kind == 0 else 65536) == 0
    ctx = L.Context(0)
    ctx.upload(data)
    if args.profile_replace:
        ctx.set_option("profile_replace", 1)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    st = None
    for i in range(args.repeat):
        st = ctx.train(args.merges)
        print(f"run {i}: ms_device {st['ms_device']:.2f} ms_total {st['ms_total']:.2f} replace_ms {st['replace_ms']:.2f} select_ms {st['select_ms']:.2f} apply_ms {st['apply_ms']:.2f} gap_ms {st['gap_ms']:.2f} passes {st['replace_passes']} batched {st['batch_merges']}", file=sys.stderr)
    out = {"train": {k: st[k] for k in ("n_input", "n_merges", "n_tokens", "kernel_launches", "replace_launches",
                                        "replace_bytes", "replace_ms", "ms_device", "table_capacity", "final_distinct")}}
    if st["replace_ms"] > 0:
        out["train"]["replace_gbs"] = st["replace_bytes"] / st["replace_ms"] / 1e6
    if args.encode:
        merges, _ = ctx.download(tokens=False)
        se = ctx.encode(merges)
        out["encode"] = {k: se[k] for k in ("n_input", "n_tokens", "ranks_applied", "kernel_launches", "ms_device")}
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
