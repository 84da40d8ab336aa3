#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: differential fuzz of the CPU oracle (all three selection modes) against the
compiled, unmodified reference (oracle/_ref/ref_harness).  Needs /root/reference to have been built
with `make -C oracle`; runs only in the development container.

usage: fuzz_vs_reference.py [n_cases] [seed] [max_len]
"""
import os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle/_ref/ref_harness")
ORA = os.path.join(ROOT, "oracle/_build/bpe_oracle_cli")


def gen(rng, max_len):
    kind = rng.integers(0, 6)
    n = int(rng.integers(2, max_len))
    if kind == 0:  # tiny alphabet: many ties, long runs
        k = int(rng.integers(1, 4))
        return rng.integers(97, 97 + k, n, dtype=np.uint8)
    if kind == 1:  # medium alphabet
        return rng.integers(32, 127, n, dtype=np.uint8)
    if kind == 2:  # full byte range (no NUL)
        return rng.integers(1, 256, n, dtype=np.uint8)
    if kind == 3:  # repeated words
        words = [rng.integers(97, 123, int(rng.integers(1, 8)), dtype=np.uint8) for _ in range(int(rng.integers(2, 40)))]
        out = []
        tot = 0
        while tot < n:
            w = words[int(rng.integers(0, len(words)))]
            out.append(w)
            out.append(np.array([32], dtype=np.uint8))
            tot += len(w) + 1
        return np.concatenate(out)[:n]
    if kind == 4:  # runs of equal bytes
        out = []
        tot = 0
        while tot < n:
            r = int(rng.integers(1, 12))
            out.append(np.full(r, rng.integers(97, 100), dtype=np.uint8))
            tot += r
        return np.concatenate(out)[:n]
    a = rng.integers(1, 256, n, dtype=np.uint8)  # with an embedded NUL
    a[int(rng.integers(0, n))] = 0
    return a


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    max_len = int(sys.argv[3]) if len(sys.argv) > 3 else 600
    rng = np.random.default_rng(seed)
    bad = 0
    with tempfile.TemporaryDirectory() as d:
        for c in range(n_cases):
            data = gen(rng, max_len)
            p = os.path.join(d, "in.bin")
            data.tofile(p)
            r = run([REF, p, "0", os.path.join(d, "ref")])
            ref_ok = r.returncode == 0
            # modes 0-2 as they are; 3 = mode 2 with helper threads and the candidate-list argmax (what the offline
            # full-size fixtures are made with)
            for mode in (0, 1, 2, 3):
                env = dict(os.environ, BO_WORKERS="3", BO_CAND_FLOOR="1") if mode == 3 else None
                o = subprocess.run([ORA, "train", p, "0", str(min(mode, 2)), os.path.join(d, f"o{mode}")], capture_output=True,
                                   text=True, env=env)
                if (o.returncode == 0) != ref_ok:
                    print(f"case {c} mode {mode}: status differs ref={r.returncode} oracle={o.returncode}")
                    bad += 1
                    continue
                if not ref_ok:
                    continue
                for ext in ("merges", "ids"):
                    a = open(os.path.join(d, f"ref.{ext}"), "rb").read()
                    b = open(os.path.join(d, f"o{mode}.{ext}"), "rb").read()
                    if a != b:
                        print(f"case {c} mode {mode}: {ext} differ (n={len(data)})")
                        data.tofile(f"/tmp/fuzz_fail_{seed}_{c}.bin")
                        bad += 1
                        break
    print(f"{n_cases} cases, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
