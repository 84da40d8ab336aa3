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
