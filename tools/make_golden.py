#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: (re)generate tests/golden/*.npz from the compiled, UNMODIFIED reference
(oracle/_ref/ref_harness, built by `make -C oracle ref` from /root/reference).  Runs only in the
development container; the fixtures it writes are committed so that the GPU box (which has no
/root/reference) can check against them.

Each fixture: input (uint8), merges ([k,2] uint32), ids (uint32), cap (0 = to exhaustion), status
(0 ok / 1 = compress() returned NULL).
"""
import os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle/_ref/ref_harness")
GOLD = os.path.join(ROOT, "tests/golden")
REFDATA = "/root/reference"

KATS = {
    "k1": b"aaaa", "k2": b"aaaaa", "k3": b"aaaaaaa", "k4": b"abababab", "k5": b"abcabcabcabc", "k6": b"a",
    "k7": b"ab", "k8": b"abab\0abababab", "k9": bytes.fromhex("fffefffefffe"), "k10": b"aabaabaab",
    "k11": b"xaaay_xaaay", "k12": b"", "k13": b"abcdabcd",
}


def run_ref(data: bytes, cap: int):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "in.bin")
        open(p, "wb").write(data)
        r = subprocess.run([REF, p, str(cap), os.path.join(d, "out")], capture_output=True, text=True)
        if r.returncode != 0:
            return 1, np.zeros((0, 2), np.uint32), np.zeros(0, np.uint32)
        m = np.fromfile(os.path.join(d, "out.merges"), dtype=np.uint32).reshape(-1, 2)
        i = np.fromfile(os.path.join(d, "out.ids"), dtype=np.uint32)
        return 0, m, i


def save(name, data, cap=0, input_ref=None):
    """input_ref: the input is a committed file under tests/golden (stored once); then only the merges, the
    number of ids and their sha256 are kept."""
    import hashlib
    status, m, i = run_ref(data, cap)
    if input_ref:
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), input_ref=np.array(input_ref), merges=m,
                            n_ids=np.int64(len(i)), ids_sha256=np.array(hashlib.sha256(i.astype("<u4").tobytes()).hexdigest()),
                            cap=np.int64(cap), status=np.int64(status))
    else:
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), input=np.frombuffer(data, dtype=np.uint8), merges=m, ids=i,
                            cap=np.int64(cap), status=np.int64(status))
    print(f"{name}: n={len(data)} cap={cap} status={status} merges={len(m)} ids={len(i)}", flush=True)


def main():
    os.makedirs(GOLD, exist_ok=True)
    which = set(sys.argv[1:])
    want = lambda n: not which or n in which
    for k, v in KATS.items():
        if want(k):
            save("kat_" + k, v)
    rt = open(os.path.join(REFDATA, "random_text.txt"), "rb").read()
    if want("testing"):
        save("testing_txt", open(os.path.join(REFDATA, "testing.txt"), "rb").read())
    if want("rt20k"):
        save("rt20k", rt[:20000])
    if want("rt64k"):
        save("rt64k", rt[:65536])
    rng = np.random.default_rng(7)
    big0 = rng.integers(1, 256, 30000, dtype=np.uint8).tobytes()
    if want("bytes30k"):
        save("bytes30k", big0)
    if want("rt_full_cap300"):  # config 1 input, first 300 merges (dynamic-regime first iteration)
        import gzip
        with gzip.GzipFile(os.path.join(GOLD, "random_text.txt.gz"), "wb", mtime=0) as g:
            g.write(rt)  # the reference's bundled data file = input of BASELINE config 1
        save("rt_full_cap300", rt, cap=300, input_ref="random_text.txt.gz")
    # exact-threshold edge: a prefix with exactly 19,661 distinct pairs (D on the doubling threshold)
    if want("edge19661"):
        a = np.random.default_rng(11).integers(1, 256, 60000, dtype=np.uint8)
        seen, L = set(), None
        for i in range(len(a) - 1):
            seen.add((int(a[i]), int(a[i + 1])))
            if len(seen) == 19661:
                L = i + 2
                break
        save("edge19661_first", a[:L].tobytes(), cap=6)
        save("edge19661_later", a[:L + 3].tobytes(), cap=6)


if __name__ == "__main__":
    main()
