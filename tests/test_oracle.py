"""CPU suite: the oracle against the committed reference fixtures, and the tiled/sharded
formulation (what the CUDA kernels and the multi-GPU protocol compute) against the plain
reference loop."""
import numpy as np
import pytest

import oracle_api
from oracle_api import FAITHFUL, FAST, FAST_CF

SMALL = [n for n in oracle_api.golden_names() if not n.startswith(("rt64k", "bytes30k", "rt_full"))]
ALL = oracle_api.golden_names()


def check_against_golden(oracle, name, mode):
    g = oracle_api.golden(name)
    rc, m, t, st = oracle.train(g["input"], g["cap"], mode)
    if g["status"]:
        assert rc != 0, "reference returned NULL (fewer than 2 characters), oracle must fail too"
        return
    assert rc == 0
    assert np.array_equal(m, g["merges"]), f"{name}: merge list differs"
    assert len(t) == g["n_ids"] and oracle_api.ids_sha(t) == g["ids_sha256"], f"{name}: ids differ"


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("mode", [FAITHFUL, FAST, FAST_CF])
def test_oracle_matches_reference_small(oracle, name, mode):
    check_against_golden(oracle, name, mode)


@pytest.mark.parametrize("name", [n for n in ALL if n not in SMALL])
@pytest.mark.parametrize("mode", [FAST, FAST_CF])
def test_oracle_matches_reference_large(oracle, name, mode):
    check_against_golden(oracle, name, mode)


@pytest.mark.parametrize("name", ["rt20k", "rt64k", "edge19661_first", "kat_k7", "testing_txt"])
def test_helper_threads_and_candidate_list_keep_the_reference_results(oracle, name):
    """The offline full-size fixtures (tools/make_full_golden.py) are made with the oracle's two whole-array loops split
    over helper threads and with the maximum taken over a candidate list; both must leave every result as it is."""
    try:
        for workers, floor in ((4, 8),) if name == "rt64k" else ((3, 1), (1, 1)):
            oracle.configure(workers=workers, min_tokens=0, candidate_floor=floor)
            for mode in (FAST_CF,) if name == "rt64k" else (FAST, FAST_CF):
                check_against_golden(oracle, name, mode)
    finally:
        oracle.configure()


def test_helper_threads_on_runs_of_one_token(oracle):
    """a == b merges: where a helper's share starts depends on the parity of the run that reaches into it.  Compared with
    the plain single-threaded whole-map scan (which the reference fixtures pin)."""
    rng = np.random.default_rng(20260)
    inputs = [np.full(4001, 97, np.uint8), np.frombuffer(b"aaab" * 2500 + b"aaaaa", dtype=np.uint8)]
    inputs += [rng.integers(97, 97 + k, size=int(rng.integers(9, 6000)), dtype=np.uint8) for k in (1, 2, 2, 3, 5)]
    try:
        for data in inputs:
            oracle.configure(workers=1, candidate_floor=0)
            rc0, m0, t0, st0 = oracle.train(data, 0, FAST_CF)
            for workers, floor in ((2, 1), (7, 3), (64, 1)):
                oracle.configure(workers=workers, min_tokens=0, candidate_floor=floor)
                rc1, m1, t1, st1 = oracle.train(data, 0, FAST_CF)
                assert rc0 == rc1 == 0 and np.array_equal(m0, m1) and np.array_equal(t0, t1) and st0 == st1
    finally:
        oracle.configure()


def test_murmur_and_bucket_counts(oracle):
    # hash_table.c:8-53 on an 8-byte key; value cross-checked with an independent Python murmur3_32
    def mm3(a, b):
        def rotl(x, r): return ((x << r) | (x >> (32 - r))) & 0xFFFFFFFF
        h = 0x9747b28c
        for k in (a, b):
            k = (k * 0xcc9e2d51) & 0xFFFFFFFF
            k = rotl(k, 15)
            k = (k * 0x1b873593) & 0xFFFFFFFF
            h ^= k
            h = rotl(h, 13)
            h = (h * 5 + 0xe6546b64) & 0xFFFFFFFF
        h ^= 8
        h ^= h >> 16
        h = (h * 0x85ebca6b) & 0xFFFFFFFF
        h ^= h >> 13
        h = (h * 0xc2b2ae35) & 0xFFFFFFFF
        h ^= h >> 16
        return h
    for a, b in [(0, 0), (97, 98), (255, 254), (256, 300), (70000, 3), (0xFFFFFFFE, 1)]:
        assert oracle.lib.bo_murmur3_pair(a, b) == mm3(a, b)
    # hash_table.c:6,248 + bpe.c:611: thresholds 19,661 / 39,322 / 78,644 / 157,287
    assert oracle.lib.bo_merged_buckets(19661) == 65536
    assert oracle.lib.bo_merged_buckets(19662) == 131072
    assert oracle.lib.bo_merged_buckets(39322) == 131072
    assert oracle.lib.bo_merged_buckets(39323) == 262144
    assert oracle.lib.bo_merged_buckets(78645) == 524288
    assert oracle.lib.bo_merged_buckets(157288) == 1048576


def test_round_trip_and_encode(oracle):
    rng = np.random.default_rng(3)
    for _ in range(40):
        n = int(rng.integers(2, 3000))
        data = rng.integers(1, int(rng.integers(3, 256)), n, dtype=np.uint8)
        rc, m, t, _ = oracle.train(data, 0, FAST_CF)
        assert rc == 0
        assert oracle.decode(t, m) == data.tobytes()           # decompress(compress(x)) == x
        assert np.array_equal(oracle.encode(data, m), t)       # encoding the training text reproduces its ids
        other = rng.integers(1, 256, 500, dtype=np.uint8)
        assert oracle.decode(oracle.encode(other, m), m) == other.tobytes()


def random_stream(rng, n, alphabet):
    kind = rng.integers(0, 3)
    if kind == 0:
        return rng.integers(0, alphabet, n).astype(np.uint32)
    if kind == 1:  # runs
        out = []
        while sum(map(len, out)) < n:
            out.append(np.full(int(rng.integers(1, 9)), rng.integers(0, alphabet), dtype=np.uint32))
        return np.concatenate(out)[:n]
    return np.full(n, rng.integers(0, alphabet), dtype=np.uint32)  # one long run


def test_sharded_tiled_rewrite_matches_reference_loop(oracle):
    """The kernel formulation (local match rule + run parity + halos + per-replacement deltas) equals the
    sequential reference loop, and its deltas equal the difference of two full recounts."""
    rng = np.random.default_rng(5)
    for it in range(1500):
        n = int(rng.integers(1, 90))
        alphabet = int(rng.integers(1, 5))
        toks = random_stream(rng, n, alphabet)
        a = int(rng.integers(0, alphabet))
        b = a if rng.random() < 0.4 else int(rng.integers(0, alphabet))
        z = alphabet + int(rng.integers(0, 3))
        ref = oracle.rewrite(toks, a, b, z)
        shards = int(rng.integers(1, 9))
        tile = int(rng.integers(1, 12))
        out, delta = oracle.rewrite_sharded(toks, a, b, z, shards, tile)
        assert np.array_equal(out, ref), (it, toks.tolist(), a, b, shards, tile)
        before, after = oracle_api.pair_counts(toks), oracle_api.pair_counts(ref)
        got = dict(before)
        def bump(x, y, d):
            k = x | (y << 32)
            got[k] = got.get(k, 0) + d
        for t in range(z + 1):
            bump(t, a, -int(delta[t, 0]))
            bump(b, t, -int(delta[t, 1]))
            bump(t, z, int(delta[t, 2]))
            bump(z, t, int(delta[t, 3]))
        got[a | (b << 32)] = 0
        got = {k: v for k, v in got.items() if v}
        assert got == after, (it, toks.tolist(), a, b, z, shards, tile)


def test_premises_of_the_batched_pass_rules(oracle):
    """The GPU engine lets a pass carry several merges when they are provably the next ones (batched
    passes).  The proof rests on three facts about ONE merge (a, b) -> z with c replacements, checked here on the
    oracle's own merge sequence (CPU only, no engine involved):
      1. only pairs that contain a, b or z change their count, and those that contain a or b never grow;
      2. a new pair never exceeds the old pair it comes from: (x, z) <= (x, a), (z, y) <= (b, y), (z, z) <= (b, a);
      3. the number of distinct pairs D grows by at most min(2c, 2V) and shrinks by at most min(2c, 2V) + 1
         (V = token ids that exist after the merge)."""
    rng = np.random.default_rng(11)
    texts = [rng.integers(97, 103, 30_000, dtype=np.uint8),                                   # few symbols, long merge chains
             np.repeat(rng.integers(97, 100, 9_000, dtype=np.uint8), rng.integers(1, 5, 9_000)),  # runs: a == b merges
             np.frombuffer((b"the quick brown fox jumps over the lazy dog " * 700), dtype=np.uint8)]
    for text in texts:
        rc, merges, ids, _ = oracle.train(text, 120, FAST)
        assert rc == 0
        toks = text.astype(np.uint32)
        before = oracle_api.pair_counts(toks)
        for r, (a, b) in enumerate(merges.tolist()):
            z = 256 + r
            c_ab = before.get(a | (b << 32), 0)
            nxt = oracle.rewrite(toks, a, b, z)
            after = oracle_api.pair_counts(nxt)
            c = (len(toks) - len(nxt))                     # replacements made (a == b: pairs of a run, <= c_ab)
            assert c <= c_ab and (a == b or c == c_ab)
            V = z + 1
            for key in set(before) | set(after):
                x, y = key & 0xFFFFFFFF, key >> 32
                old, new = before.get(key, 0), after.get(key, 0)
                if old != new:
                    assert {x, y} & {a, b, z}, (r, x, y)                               # 1: nothing else moves
                if z not in (x, y):
                    assert new <= old, (r, x, y, old, new)                             # 1: old pairs never grow
                elif x == z and y == z:
                    assert new <= before.get(b | (a << 32), 0)                         # 2
                elif y == z:
                    assert new <= before.get(x | (a << 32), 0), (r, x, new)            # 2
                else:
                    assert new <= before.get(b | (y << 32), 0), (r, y, new)            # 2
            bound = min(2 * c, 2 * V)
            assert len(after) - len(before) <= bound and len(before) - len(after) <= bound + 1, (r, len(before), len(after), c)
            toks, before = nxt, after
        assert np.array_equal(toks, ids)


@pytest.mark.parametrize("kind", ["zipf_words", "zipf_bytes", "near_uniform"])
def test_batch_walk_rules_predict_the_oracles_next_merges(oracle, kind):
    """The rules by which apply_select_kernel lets a pass carry several merges (disjoint tokens, no
    same-bucket tie, strictly above the bound - or equal to it and untouched -, D clear of every doubling threshold),
    restated here in Python and run against the oracle's own merge order: every batch they form must be exactly the
    oracle's next merges, in order.  (The CUDA implementation of the same rules is checked by the -m gpu tests.)"""
    from llmtokenizer_b200 import _lib
    if kind in ("zipf_words", "zipf_bytes"):
        text = np.zeros(400_000, dtype=np.uint8)
        assert _lib.load_corpus().gen_corpus_fill(0 if kind == "zipf_words" else 1, text.ctypes.data, text.size, 5,
                                                  50000 if kind == "zipf_words" else 65536) == 0
        text = text[:len(text.tobytes().split(b"\0")[0])]
        cap = 500
    else:   # 40 almost equally frequent symbols: late counts tie all the time, the equal-count rule decides
        text = np.random.default_rng(3).integers(40, 80, 300_000, dtype=np.uint8)
        cap = 700
    rc, merges, ids, _ = oracle.train(text, cap, FAST_CF)
    assert rc == 0 and len(merges) == cap
    merges = [tuple(m) for m in merges.tolist()]
    mm3, buckets = oracle.lib.bo_murmur3_pair, oracle.lib.bo_merged_buckets
    toks = text.astype(np.uint32)
    r, batches, carried = 0, 0, 0
    while r < cap:
        counts = oracle_api.pair_counts(toks)
        D = len(counts)
        B = buckets(D)
        order = sorted(((cnt, mm3(k & 0xFFFFFFFF, k >> 32) % B, k) for k, cnt in counts.items() if cnt >= 2),
                       key=lambda e: (-e[0], e[1]))[:200]
        assert (order[0][2] & 0xFFFFFFFF, order[0][2] >> 32) == merges[r] or order[0][:2] == order[1][:2]
        acc = [merges[r]]                           # what decide() committed; the walk extends it
        acc_cnt = [counts[merges[r][0] | (merges[r][1] << 32)]]
        bound, rescue_ok = 0, True
        if merges[r][0] != merges[r][1] and order[0][:2] != order[1][:2]:
            for i in range(1, len(order)):
                cnt, bkt, k = order[i]
                a, b = k & 0xFFFFFFFF, k >> 32
                tie = (i + 1 < len(order) and order[i + 1][:2] == (cnt, bkt)) or order[i - 1][:2] == (cnt, bkt)
                overlap = any(t in (a, b) for p in acc for t in p)
                if tie or a == b or overlap or len(acc) >= 8:
                    bound, rescue_ok = cnt, not tie
                    break
                acc.append((a, b))
                acc_cnt.append(cnt)
            else:
                bound = order[-1][0]                # (the walk never gets this far: 200 candidates, 8 merges)
            n = sum(1 for c in acc_cnt if c > bound)
            if rescue_ok and n < len(acc):
                # merges whose count equals the bound survive unless a pair of exactly that count touches a merge in front
                jmin = len(acc)
                for cnt, _, k in order:
                    a, b = k & 0xFFFFFFFF, k >> 32
                    if cnt == bound and (a, b) not in acc:
                        for j, p in enumerate(acc):
                            if a in p or b in p:
                                jmin = min(jmin, j)
                                break
                n = max(n, min(len(acc), jmin + 1))
            n = max(n, 1)
            # D may move by min(2c, 2V) (+1 downwards) per merge in front: B(D) must not be able to change
            V, up, down = 256 + r + 8, 0, 0
            for j in range(1, n):
                up += min(2 * acc_cnt[j - 1], 2 * V)
                down += min(2 * acc_cnt[j - 1], 2 * V) + 1
                if buckets(max(D - down - 1, 0)) != buckets(D + up + 1):
                    n = j
                    break
            acc = acc[:n]
        else:
            acc = acc[:1]
        acc = acc[:cap - r]                         # (the engine's `room`: the merge cap ends a batch)
        assert acc == merges[r:r + len(acc)], (r, acc, merges[r:r + len(acc)])
        batches += 1
        carried += len(acc) - 1
        for a, b in acc:
            toks = oracle.rewrite(toks, a, b, 256 + r)
            r += 1
            if r >= cap:
                break
    assert np.array_equal(toks, ids)
    assert carried > cap // 8, (batches, carried)       # the rules are not vacuous: many merges ride along
