"""GPU suite (-m gpu): the CUDA engine, called through its C ABI, against the committed reference
fixtures and the CPU oracle on the same seeded inputs.  Bit-exact: merge list, ids.

Small inputs run the oracle inline; the heavier seeded cases compare with the oracle's committed results
(tests/parity_cases.py, tests/golden/cases/, made offline by tools/make_case_digests.py) so that the whole
suite stays inside the GPU box's time budget."""
import numpy as np
import pytest

import oracle_api
import parity_cases as pc
from oracle_api import FAST, FAST_CF
from parity_cases import corpus, sha

pytestmark = pytest.mark.gpu


def expected_inline(oracle, data, cap=0):
    rc, om, ot, ost = oracle.train(data, cap, FAST_CF)
    assert rc == 0
    return {"merges": om, "n_ids": len(ot), "ids_sha256": sha(ot), "same_bucket_ties": ost["same_bucket_ties"],
            "threshold_edges": ost["threshold_edges"], "thread_buckets": ost["thread_buckets"]}


def check_result(exp, m, t, st, n_gpus=1, what=""):
    om = exp["merges"]
    k = min(len(m), len(om))
    first_bad = next((i for i in range(k) if tuple(m[i]) != tuple(om[i])), None)
    assert first_bad is None and len(m) == len(om), (
        f"{what}: merge lists differ (engine {len(m)} merges, oracle {len(om)}, first mismatch at {first_bad}: "
        f"engine {m[first_bad].tolist() if first_bad is not None else None} oracle "
        f"{om[first_bad].tolist() if first_bad is not None else None}); engine stats {st}")
    assert len(t) == exp["n_ids"] and sha(t) == exp["ids_sha256"], f"{what}: ids differ; engine stats {st}"
    if True:
        # the emulated bucket counts of the reference's 16 worker tables must track the oracle's exactly (any number of GPUs)
        assert st["worker_buckets"] == list(exp["thread_buckets"]), (what, st["worker_buckets"], exp["thread_buckets"])
        assert st["same_bucket_ties"] == exp["same_bucket_ties"] and st["threshold_edges"] == exp["threshold_edges"], (what, st)


def assert_same(engine, oracle, data, cap=0, n_gpus=1, what=""):
    """engine == oracle (run inline) on `data`"""
    exp = expected_inline(oracle, data, cap)
    m, t, st = engine.train(data, max_merges=cap, n_gpus=n_gpus)
    check_result(exp, m, t, st, n_gpus, what)
    return m, t, st


def assert_case(engine, oracle, key, n_gpus=1, options=None):
    """engine == the oracle's committed result for parity_cases.TRAIN_CASES[key]"""
    data, cap = pc.train_input(key)
    exp = pc.expected_train(key, oracle)
    if options:
        ctx = engine.Context(0)
        try:
            for k, v in options.items():
                ctx.set_option(k, v)
            ctx.upload(data)
            st = ctx.train(cap)
            m, t = ctx.download()
        finally:
            ctx.close()
    else:
        m, t, st = engine.train(data, max_merges=cap, n_gpus=n_gpus)
    check_result(exp, m, t, st, n_gpus, f"{key} P={n_gpus} {options or ''}")
    return data, m, t, st


def assert_encode_case(engine, oracle, key, m, n_gpus=1):
    exp = pc.expected_encode(key, oracle)
    ids, st = engine.encode(pc.ENCODE_CASES[key][1](), m, n_gpus=n_gpus)
    assert len(ids) == exp["n_ids"] and sha(ids) == exp["ids_sha256"], (key, st)


# ---- regressions first: inputs whose candidate list runs empty (round 1: the host loop span forever) ---------
@pytest.mark.parametrize("n", [31, 62, 3839, 4000])
def test_single_run_to_exhaustion(engine, oracle, n):
    """b"a" * n: every merge halves the best count (3999, 1999, ... 30, 14), so a candidate list built for
    "at least half the maximum" is empty after each merge (bpe.c:737-750 picks the new maximum every time)."""
    assert_same(engine, oracle, np.full(n, 97, dtype=np.uint8), what=f"a*{n}")


@pytest.mark.parametrize("symbols", [2, 3])
def test_tiny_alphabets_to_exhaustion(engine, oracle, symbols):
    rng = np.random.default_rng(100 + symbols)
    for n in (40, 700, 3999, 20_000):
        assert_same(engine, oracle, rng.integers(97, 97 + symbols, n, dtype=np.uint8), what=f"{symbols} symbols n={n}")
    # long runs with rare breaks: counts fall steeply, lists are used up again and again
    data = np.repeat(rng.integers(97, 97 + symbols, 400, dtype=np.uint8), rng.integers(1, 200, 400))
    assert_same(engine, oracle, data, what=f"{symbols} symbols, long runs")


def test_steep_counts_above_the_static_limit(engine, oracle):
    """The same shapes in the streaming regime (>= 1,048,576 tokens): one long run, and a byte-level Zipf text
    whose first merges have a handful of pairs above half the maximum (lists that run empty, thresholds that
    are lowered a quarter at a time)."""
    assert_same(engine, oracle, np.full(2_500_000, 97, dtype=np.uint8), what="a*2.5M")
    assert_case(engine, oracle, "zipfb3m_1500")
    assert_case(engine, oracle, "zipfb3m_1500", options={"batch_max": 1})


def test_crossing_the_static_limit_inside_batched_passes(engine, oracle):
    """1.4 M tokens that fall below 1,048,576 (where the reference starts slicing statically, bpe.c:449) while
    merges share passes: the crossing iteration must be selected on its own so that the 16 worker tables' bucket
    counts (compared by check_result) follow the oracle's.  The low histogram limit turns batching on from id 260."""
    assert_case(engine, oracle, "cross1m_400", options={"smem_hist_max_vocab": 260, "batch_min_z": 260})
    assert_case(engine, oracle, "cross1m_400")


def test_ties_above_the_static_limit(engine, oracle):
    assert_case(engine, oracle, "ties4m_1000")


def test_tie_heavy_text_to_exhaustion(engine, oracle):
    """60 shifted copies of a 20 KB random text (1.2 M tokens), to exhaustion: 15,104 merges, 419 of them decided
    by the chain order inside one bucket (hash_table.c:208-223,300-302), above and below the static limit."""
    _, _, _, st = assert_case(engine, oracle, "ties1m2_exh")
    assert st["same_bucket_ties"] == 419


@pytest.mark.parametrize("name", oracle_api.golden_names())
def test_reference_fixtures(engine, name):
    g = oracle_api.golden(name)
    if g["status"]:
        with pytest.raises(engine.BpeCudaError) as e:
            engine.train(g["input"], max_merges=g["cap"])
        assert e.value.rc == -2      # "File contains less than 2 characters" (bpe.c:558-563)
        return
    m, t, st = engine.train(g["input"], max_merges=g["cap"])
    k = min(len(m), len(g["merges"]))
    first_bad = next((i for i in range(k) if tuple(m[i]) != tuple(g["merges"][i])), None)
    assert first_bad is None and len(m) == len(g["merges"]), (name, len(m), len(g["merges"]), first_bad, st)
    assert len(t) == g["n_ids"] and oracle_api.ids_sha(t) == g["ids_sha256"], (name, st)


def test_random_small_inputs(engine, oracle):
    rng = np.random.default_rng(17)
    for case in range(60):
        n = int(rng.integers(2, 4000))
        kind = case % 4
        if kind == 0:
            data = rng.integers(97, 97 + int(rng.integers(1, 4)), n, dtype=np.uint8)
        elif kind == 1:
            data = rng.integers(32, 127, n, dtype=np.uint8)
        elif kind == 2:
            data = rng.integers(1, 256, n, dtype=np.uint8)
        else:
            data = np.repeat(rng.integers(97, 101, n // 3 + 1, dtype=np.uint8), rng.integers(1, 7, n // 3 + 1))[:n]
        assert_same(engine, oracle, data, what=f"case {case} kind {kind} n {n}")


@pytest.mark.parametrize("n", [3839, 3840, 3841, 7679, 7680, 7681, 3840 * 5 + 1, 3840 * 40 - 2])
def test_tile_boundaries_and_runs(engine, oracle, n):
    # tile size of the replace kernel is 3,840 tokens: streams that end on / around tile edges,
    # made of long runs so that a == b merges and their run parity cross tiles
    rng = np.random.default_rng(n)
    runs = np.repeat(rng.integers(97, 100, n, dtype=np.uint8), rng.integers(1, 40, n))[:n]
    assert_same(engine, oracle, runs, cap=40, what=f"runs n={n}")
    allsame = np.full(n, 97, dtype=np.uint8)
    assert_same(engine, oracle, allsame, what=f"all-a n={n}")
    ab = np.tile(np.frombuffer(b"ab", dtype=np.uint8), n // 2 + 1)[:n]
    assert_same(engine, oracle, ab, what=f"abab n={n}")


def test_medium_corpora_capped(engine, oracle):
    assert_case(engine, oracle, "zipfa6m_300")      # zipf_ascii 6 MB / 300 merges
    assert_case(engine, oracle, "zipfb5m_150")      # zipf_bytes 5 MB / 150 merges
    assert_case(engine, oracle, "rt2p4m_120")       # 2.4 MB random text, dynamic regime throughout / 120 merges


@pytest.mark.parametrize("ranges", [1, 5, 64, 0])
def test_ranged_stream_boundaries(engine, oracle, ranges):
    """The streaming kernel keeps the stream as independently compacted ranges; whatever their number
    and however their lengths fall relative to the 4,096-token tiles, results must not change.  Small
    alphabets make replacements land on range ends all the time (including ranges whose last tile holds
    one or two tokens, and a == b merges that force a repack)."""
    rng = np.random.default_rng(1000 + ranges)
    cases = [rng.integers(97, 100, 1_300_000, dtype=np.uint8),
             np.repeat(rng.integers(97, 101, 500_000, dtype=np.uint8), rng.integers(1, 5, 500_000))[:1_200_000],
             corpus(0, 1_500_000, 31 + ranges)]
    for i, data in enumerate(cases):
        rc, om, ot, _ = oracle.train(data, 60, FAST_CF)
        ctx = engine.Context(0)
        ctx.set_option("ranges", ranges)
        ctx.upload(data)
        ctx.train(60)
        m, t = ctx.download()
        assert np.array_equal(m, om) and np.array_equal(t, ot), f"train case {i} ranges {ranges}"
        st = ctx.encode(om)
        _, t2 = ctx.download()
        ctx.close()
        assert np.array_equal(t2, ot), f"encode case {i} ranges {ranges}: {st}"


def test_batched_passes_long_runs(engine, oracle):
    """Late in training most passes carry several provably-next merges (batched passes).  Long
    capped runs on corpora that stay above 1,048,576 tokens: the merge ORDER (ids) must be the sequential one."""
    for key, kind in (("zipfa12m_2500", 0), ("zipfb8m_2200", 1), ("zipfa9m_2000", 0)):
        data, m, t, st = assert_case(engine, oracle, key)
        cap = len(m)
        if kind == 0:   # (merges share passes once the ids have outgrown the shared-memory delta histogram)
            assert st["batch_merges"] > 0 and st["replace_passes"] < cap, st
        # and with batching off the very same result
        ctx = engine.Context(0)
        ctx.set_option("batch_max", 1)
        ctx.upload(data)
        s1 = ctx.train(cap)
        m1, t1 = ctx.download()
        ctx.close()
        assert s1["batch_merges"] == 0 and np.array_equal(m1, m) and np.array_equal(t1, t)
        # encoding batches consecutive ranks with unconnected tokens: same ids as training, and as the oracle elsewhere
        ids, se = engine.encode(data, m)
        assert np.array_equal(ids, t), se
        assert_encode_case(engine, oracle, "enc_" + key, m)


def test_batched_passes_beyond_the_class_table(engine, oracle):
    """Batched passes find a token's pair through an 8,192-entry table indexed by (id mod 8192): once ids
    pass 8,192 they alias table entries, candidates are checked against the real pair, and pairs whose tokens
    alias each other may not share a pass.  10,000 merges on a stream that stays above 1,048,576 tokens;
    checked against the unbatched engine and against the oracle's committed digests; the oracle also checks
    the encoder on a smaller text and the decoder closes the loop."""
    key = "zipfa12m_10000"
    data, cap = pc.train_input(key)
    exp = pc.expected_train(key, oracle)
    res = []
    for bm in (12, 1):
        ctx = engine.Context(0)
        ctx.set_option("batch_max", bm)
        ctx.upload(data)
        st = ctx.train(cap)
        m, t = ctx.download()
        if bm == 12:
            assert st["batch_merges"] > 0 and ctx.decode(m, download=False) == data.size and ctx.decode_mismatches() == 0
        ctx.close()
        check_result(exp, m, t, st, 1, f"{key} batch_max {bm}")
        res.append((m, t, st))
    (m, t, st), (m1, t1, s1) = res
    assert len(m) == cap and s1["batch_merges"] == 0
    ids, se = engine.encode(data, m)
    assert np.array_equal(ids, t) and se["batch_merges"] > 0, se
    assert_encode_case(engine, oracle, "enc_" + key, m)


def test_batched_passes_with_many_ties(engine, oracle):
    """Nearly uniform symbols: late merges all have almost the same count, so the order inside and between
    batches is decided by the bucket order again and again (and now and then by a same-bucket tie, which
    must end a batch and go through the exact resolver)."""
    _, m, t, st = assert_case(engine, oracle, "uniform14m_2600")
    assert st["replace_passes"] <= 2600


def test_encode_matches_training_ids_and_oracle(engine, oracle):
    data = corpus(0, 1_500_000, 5)
    m, t, _ = engine.train(data, max_merges=400)
    ids, st = engine.encode(data, m)
    assert np.array_equal(ids, t), "encoding the training text with its own merges must reproduce compress()'s ids"
    other = corpus(0, 700_000, 6)
    ids2, st2 = engine.encode(other, m)
    assert np.array_equal(ids2, oracle.encode(other, m))
    assert oracle.decode(ids2, m) == other.tobytes()
    assert st2["ranks_applied"] <= 400


def test_nul_truncation_and_unsigned_bytes(engine, oracle):
    data = np.frombuffer(b"abab\0abababab", dtype=np.uint8)
    m, t, _ = engine.train(data)
    assert m.tolist() == [[97, 98]] and t.tolist() == [256, 256]
    hi = np.frombuffer(bytes.fromhex("fffefffefffe"), dtype=np.uint8)
    m, t, _ = engine.train(hi)
    assert m.tolist() == [[255, 254], [256, 256]] and t.tolist() == [257, 256]  # k9


def test_repeated_calls_and_context_reuse(engine, oracle):
    # the reference's compress() can be called once per process; the engine must not care
    data = corpus(0, 300_000, 1)
    ctx = engine.Context(0)
    ctx.upload(data)
    outs = []
    for _ in range(3):
        ctx.train(50)
        outs.append(ctx.download())
    ctx.close()
    for m, t in outs[1:]:
        assert np.array_equal(m, outs[0][0]) and np.array_equal(t, outs[0][1])
    rc, om, ot, _ = oracle.train(data, 50, FAST)
    assert np.array_equal(outs[0][0], om) and np.array_equal(outs[0][1], ot)


def _device_count():
    from llmtokenizer_b200 import _lib
    return _lib.load().bpe_cuda_device_count()


@pytest.mark.parametrize("P", [2, 4, 8])
def test_sharded_training_and_encoding_match_oracle(engine, oracle, P):
    """Corpus sharded over P GPUs (one NCCL all-reduce of the delta vectors + edge records per merge) must give
    exactly the single-stream result: merges, ids, any P."""
    if _device_count() < P:
        pytest.skip(f"needs {P} GPUs")
    assert_case(engine, oracle, "zipfa8m_200", n_gpus=P)
    assert_case(engine, oracle, "zipfb6m_100", n_gpus=P)
    # long enough for merges to share passes (ids beyond the shared-memory histogram), all ranks deciding alike
    _, _, _, st = assert_case(engine, oracle, "zipfa12m_1300", n_gpus=P)
    assert st["batch_merges"] > 0, st
    # long runs of equal bytes: a == b merges whose run parity crosses shard boundaries
    assert_case(engine, oracle, f"runs4m_12_P{P}", n_gpus=P)
    assert_case(engine, oracle, "same3m_1", n_gpus=P)
    data, m, t, _ = assert_case(engine, oracle, "zipfa3m_300")
    ids, _ = engine.encode(data, m, n_gpus=P)
    assert np.array_equal(ids, t)
    assert_encode_case(engine, oracle, "enc_zipfa3m_300", m, n_gpus=P)


@pytest.mark.parametrize("P", [2, 8])
def test_sharded_runs_are_exact_including_the_tie_break(engine, oracle, P):
    """Chain order inside one bucket (hash_table.c:208-223,300-302) and the 16 static slices (bpe.c:449-477) depend
    on positions in the WHOLE stream: below 1,048,576 tokens a sharded run is consolidated on every rank, above it a
    tie gathers the stream for the resolver.  Same merges, ids, tie counters and worker-table bucket counts as the
    oracle - i.e. as one GPU."""
    if _device_count() < P:
        pytest.skip(f"needs {P} GPUs")
    # the reference fixtures (all below the static limit: consolidated at once; rt20k has 9 same-bucket ties,
    # edge19661_* sit on a doubling threshold, the KATs leave most ranks with an empty shard)
    for name in oracle_api.golden_names():
        g = oracle_api.golden(name)
        if g["status"] or name.startswith("rt_full"):
            continue
        m, t, st = engine.train(g["input"], max_merges=g["cap"], n_gpus=P)
        assert np.array_equal(m, g["merges"]) and len(t) == g["n_ids"] and oracle_api.ids_sha(t) == g["ids_sha256"], (name, P, st)
    # 1.2 M tokens to exhaustion, 419 same-bucket ties: sharded at first, consolidated when the stream has shrunk
    _, _, _, st = assert_case(engine, oracle, "ties1m2_exh", n_gpus=P)
    assert st["same_bucket_ties"] == 419
    # stays above the static limit for all 1,000 merges, 4 same-bucket ties (the stream is gathered for the resolver)
    _, _, _, st = assert_case(engine, oracle, "ties4m_1000", n_gpus=P)
    assert st["same_bucket_ties"] == 4 and st["resolver_runs"] >= 4
    # falls below the limit in the middle of batched passes
    assert_case(engine, oracle, "cross1m_400", n_gpus=P)


@pytest.mark.parametrize("P", [2, 8])
def test_exchange_inbox_growth(engine, oracle, P, monkeypatch):
    """The peers' inboxes start at 64 entries per slot here (default 262,144): they must be re-allocated and their
    handles re-exchanged in mid-run, several times, without anybody noticing."""
    if _device_count() < P:
        pytest.skip(f"needs {P} GPUs")
    monkeypatch.setenv("BPE_CUDA_XCHG_CAP", "64")
    assert_case(engine, oracle, "zipfa8m_200", n_gpus=P)
    assert_case(engine, oracle, "zipfa12m_1300", n_gpus=P)


def test_more_gpus_than_devices_is_an_error(engine):
    with pytest.raises(engine.BpeCudaError) as e:
        engine.train(np.frombuffer(b"abababab" * 100, dtype=np.uint8), n_gpus=_device_count() + 1)
    assert e.value.rc in (-4, -1)


def test_file_ingest_equals_training_from_memory(engine, tmp_path):
    """bpe_cuda_train_file (bpe.c:130-180 get_file + :555 strlen cut + :580-584 widen in front of the path): the file is
    read in 32 MB pinned pieces that are copied, widened and counted while the next piece is read.  Same merges and ids
    as the in-memory entry point: pairs that straddle two pieces, a 0x00 in the third piece, files shorter than a piece,
    and the same through the drop-in's compress() a reference caller uses."""
    data = corpus(0, 75_000_000, 41)                 # three pieces: 32 + 32 + 11 MB
    data[70_000_001] = 0                             # strlen cut inside the third piece
    data[33_554_431], data[33_554_432] = 113, 117    # a pair across the first piece boundary
    p = tmp_path / "corpus.bin"
    data.tofile(p)
    m0, t0, _ = engine.train(data, max_merges=300)
    m1, t1, st = engine.train_file(str(p), max_merges=300)
    assert st["n_input"] == 70_000_001 and np.array_equal(m0, m1) and np.array_equal(t0, t1)
    ids0, _ = engine.encode(data, m0)
    ids1, _ = engine.encode_file(str(p), m0)
    assert np.array_equal(ids0, ids1) and np.array_equal(ids0, t0)
    small = tmp_path / "small.bin"
    data[:1_000_003].tofile(small)
    ms, ts, _ = engine.train(data[:1_000_003], max_merges=50)
    mf, tf, _ = engine.train_file(str(small), max_merges=50)
    assert np.array_equal(ms, mf) and np.array_equal(ts, tf)
    pair_arr, ids = engine.compress(str(p), max_merges=300)
    assert np.array_equal(pair_arr[256:], m0) and np.array_equal(ids, t0)
    if _device_count() >= 2:                          # every rank reads its own byte range; the NUL is rank 1's
        m2, t2, st2 = engine.train_file(str(p), max_merges=300, n_gpus=2)
        assert np.array_equal(m0, m2) and np.array_equal(t0, t2)
        data[20_000_000] = 0                          # ... and now rank 0's: rank 1's part is dropped
        data.tofile(p)
        m3, t3, _ = engine.train(data, max_merges=100)
        m4, t4, _ = engine.train_file(str(p), max_merges=100, n_gpus=2)
        assert np.array_equal(m3, m4) and np.array_equal(t3, t4)


def _full_golden(name):
    import json
    import os
    here = os.path.dirname(__file__)
    g = json.load(open(os.path.join(here, "golden", name + ".json")))
    g["merge_list"] = np.load(os.path.join(here, "golden", "full", name + "_merges.npz"))["merges"]
    return g


def test_config1_random_text_to_exhaustion_matches_the_oracle(engine):
    """BASELINE config 1: the reference's bundled random_text.txt (1,048,576 bytes) with its default merge count,
    i.e. to exhaustion (bpe.c:745): 39,716 merges, 154 of them same-bucket ties, iteration 0 in the dynamic regime
    (n = 2^20 is not below the limit, bpe.c:449), all others in the static one.  The oracle needs 11 minutes; its merge
    list and the digest of its ids are committed (tools/make_full_golden.py c1_exhaustion)."""
    g = _full_golden("c1_exhaustion")
    data = oracle_api.golden("rt_full_cap300")["input"]
    assert data.size == g["corpus"]["bytes"]
    m, t, st = engine.train(data)
    exp = {"merges": g["merge_list"], "n_ids": g["n_ids"], "ids_sha256": g["ids_sha256"], "same_bucket_ties": g["same_bucket_ties"],
           "threshold_edges": g["threshold_edges"], "thread_buckets": g["thread_buckets"]}
    check_result(exp, m, t, st, 1, "config 1 to exhaustion")
    assert len(m) == 39716 and st["final_distinct"] == g["final_distinct"]


def test_config2_full_size_matches_the_oracle_digest(engine):
    """BASELINE config 2 at its full size (100 MB, 4,096 merges): the oracle needs four minutes for it, so its
    result is committed as two SHA-256 digests (tests/golden/c2_full.json, tools/make_c2_golden.py)."""
    import hashlib
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c2_full.json")))
    data = corpus(0, g["corpus"]["bytes"], g["corpus"]["seed"])
    m, t, st = engine.train(data, max_merges=g["merges"])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype="<u4").tobytes()).hexdigest()
    assert len(t) == g["n_ids"] and sha(m) == g["merges_sha256"] and sha(t) == g["ids_sha256"], st
    assert st["batch_merges"] > 0


def test_config3_first_4000_merges_match_the_oracle_digest(engine):
    """BASELINE config 3's corpus at full size (1 GB of Zipf bytes, 296 full-size ranges, 32-bit counts in the
    hundreds of millions, several table rehashes, batched passes): its first 4,000 merges cost the oracle an hour,
    so they are committed as digests (tests/golden/c3_first4000.json, tools/make_c2_golden.py)."""
    import hashlib
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c3_first4000.json")))
    data = corpus(1, g["corpus"]["bytes"], g["corpus"]["seed"])
    ctx = engine.Context(0)
    try:
        ctx.upload(data)
        st = ctx.train(g["merges"])
        m, t = ctx.download()
        assert ctx.decode(m, download=False) == data.size and ctx.decode_mismatches() == 0
    finally:
        ctx.close()
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype="<u4").tobytes()).hexdigest()
    assert len(t) == g["n_ids"] and sha(m) == g["merges_sha256"] and sha(t) == g["ids_sha256"], st
    assert st["batch_merges"] > 0


def test_config3_all_32000_merges_match_the_oracle(engine):
    """BASELINE config 3 in full: 1 GB of Zipf bytes, 32,000 merges.  The oracle's merge list and the digest of its
    293 M ids are committed (tests/golden/c3_full.json + full/c3_full_merges.npz, made offline by
    tools/make_full_golden.py with the oracle's helper threads); every pass of the run is behind them - the table grows
    to tens of millions of pairs through several rehashes, ids pass the 8,192-entry class table four times over."""
    g = _full_golden("c3_full")
    data = corpus(1, g["corpus"]["bytes"], g["corpus"]["seed"])
    ctx = engine.Context(0)
    try:
        ctx.upload(data)
        st = ctx.train(g["cap"])
        m, t = ctx.download()
    finally:
        ctx.close()
    exp = {"merges": g["merge_list"], "n_ids": g["n_ids"], "ids_sha256": g["ids_sha256"], "same_bucket_ties": g["same_bucket_ties"],
           "threshold_edges": g["threshold_edges"], "thread_buckets": g["thread_buckets"]}
    check_result(exp, m, t, st, 1, "config 3, all 32,000 merges")
    assert st["final_distinct"] == g["final_distinct"] and st["batch_merges"] > 0


def test_config5_first_chunk_50000_merges_match_the_oracle(engine):
    """BASELINE config 5's vocabulary size on the first 125 MB chunk of its corpus (zipf_bytes seed 8888): 50,000 merges
    (GPT-2 scale: ids up to 50,255, delta vectors of 12 x 4 x 50 k counters, the class table aliased six times over).
    The 8 GB corpus itself is beyond the oracle's memory; its merge list for the chunk is committed
    (tests/golden/c5_chunk0_50k.json + full/, 16 minutes of oracle time)."""
    g = _full_golden("c5_chunk0_50k")
    data = corpus(1, g["corpus"]["bytes"], g["corpus"]["seed"])
    m, t, st = engine.train(data, max_merges=g["cap"])
    exp = {"merges": g["merge_list"], "n_ids": g["n_ids"], "ids_sha256": g["ids_sha256"], "same_bucket_ties": g["same_bucket_ties"],
           "threshold_edges": g["threshold_edges"], "thread_buckets": g["thread_buckets"]}
    check_result(exp, m, t, st, 1, "config 5, first chunk, 50,000 merges")
    assert st["final_distinct"] == g["final_distinct"] and st["batch_merges"] > 0
    enc, _ = engine.encode(data, m)                   # the 50,000-rank table through the encoder: the training ids again
    assert len(enc) == g["n_ids"] and sha(enc) == g["ids_sha256"]


# ---- decode: ids -> bytes, the inverse of the path ------------------------
@pytest.mark.parametrize("kind,size,cap", [(0, 300_000, 600), (1, 200_000, 300), (2, 150_000, 200)])
def test_decode_matches_oracle_and_round_trips(engine, oracle, kind, size, cap):
    data = corpus(kind, size, 7 + kind)
    m, t, _ = engine.train(data, max_merges=cap)
    got, st = engine.decode(t, m)
    assert got == oracle.decode(t, m)
    cut = data.tobytes().split(b"\0")[0]
    assert got == cut                      # decode(train(x).ids) == x
    assert st["n_tokens"] == len(cut) and st["kernel_launches"] == 3
    # a vocabulary prefix decodes ids produced by the same prefix
    ids, _ = engine.encode(data, m[: cap // 2])
    assert engine.decode(ids, m[: cap // 2])[0] == cut


def test_decode_edge_cases(engine, oracle):
    none = np.zeros((0, 2), np.uint32)
    assert engine.decode(np.zeros(0, np.uint32), none)[0] == b""
    raw = np.arange(256, dtype=np.uint32)          # every byte value, 0x00 included: NUL-safe
    assert engine.decode(raw, none)[0] == bytes(range(256))
    # a == b chains double the expansion: id 256+k is 2^(k+1) bytes, far beyond one staging tile
    m = np.array([[97, 97]] + [[256 + k, 256 + k] for k in range(17)], dtype=np.uint32)
    ids = np.array([273, 98, 256, 273, 0, 260], dtype=np.uint32)
    got = engine.decode(ids, m)[0]
    assert got == oracle.decode(ids, m) and len(got) == 2 * (1 << 18) + 1 + 2 + 1 + 32
    with pytest.raises(engine.BpeCudaError) as e:   # id outside the vocabulary
        engine.decode(np.array([97, 300], np.uint32), m[:3])
    assert e.value.rc == -1
    with pytest.raises(engine.BpeCudaError) as e:   # merge referring to a later id
        engine.decode(np.array([97], np.uint32), np.array([[300, 97]], np.uint32))
    assert e.value.rc == -1
    # ragged sizes around the 2,048-id tile
    rng = np.random.default_rng(5)
    mm = np.array([[97, 98], [256, 99], [257, 257], [258, 100]], dtype=np.uint32)
    for n in (1, 2047, 2048, 2049, 4097, 50_001):
        ids = rng.integers(0, 260, n).astype(np.uint32)
        assert engine.decode(ids, mm)[0] == oracle.decode(ids, mm), n


def test_decode_on_device_round_trip(engine):
    # train, then decode the resident stream and compare it with the resident shard without any host copy
    data = corpus(0, 4_000_000, 3)
    data = data[: len(data.tobytes().split(b"\0")[0])]
    ctx = engine.Context(0)
    try:
        ctx.upload(data)
        ctx.train(500)
        m, t = ctx.download()
        assert ctx.decode(m, download=False) == data.size
        assert ctx.decode_mismatches() == 0
        assert ctx.decode(m) == data.tobytes()
        # encode with the learned table on the same context, decode again
        ctx.encode(m)
        assert ctx.decode(m, download=False) == data.size and ctx.decode_mismatches() == 0
        # a wrong table must be noticed by the comparison
        bad = m.copy()
        bad[0] = bad[0][::-1] if bad[0][0] != bad[0][1] else [bad[0][0], bad[0][0] ^ 1]
        assert ctx.decode(bad, download=False) == data.size
        assert ctx.decode_mismatches() > 0
    finally:
        ctx.close()


def test_zipf_bytes_200mb_10000_merges_match_the_oracle_digest(engine):
    """Byte-level Zipf corpus, 200 MB, 10,000 merges: batched passes with ids beyond the 8,192-entry class table
    on a stream of 80+ M tokens and a table of millions of pairs (33 minutes of oracle time, committed as digests)."""
    import hashlib
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "zipfb200m_10k.json")))
    data = corpus(1, g["corpus"]["bytes"], g["corpus"]["seed"])
    m, t, st = engine.train(data, max_merges=g["merges"])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype="<u4").tobytes()).hexdigest()
    assert len(t) == g["n_ids"] and sha(m) == g["merges_sha256"] and sha(t) == g["ids_sha256"], st
    assert st["batch_merges"] > 0


def test_config4_first_chunk_encoded_with_config3s_table(engine):
    """BASELINE config 4's shape: a chunk of ITS corpus (zipf_bytes sampling seed 777, config 3's word list) encoded with
    the 32,000-merge table learned from config 3's corpus - a foreign table, so some ranks never occur.  The oracle's
    rank-by-rank rewrite (bpe.c:760-772 per rank) of the 125 MB chunk took 34 minutes; the digest of its 36.7 M ids is
    committed (tests/golden/c4_chunk0_encode.json, tools/make_encode_golden.py)."""
    import json
    import os
    here = os.path.dirname(__file__)
    g = json.load(open(os.path.join(here, "golden", "c4_chunk0_encode.json")))
    merges = np.load(os.path.join(here, "golden", "full", "c3_full_merges.npz"))["merges"]
    assert len(merges) == g["ranks"]
    data = corpus(1, g["corpus"]["bytes"], g["corpus"]["seed"])
    ids, st = engine.encode(data, merges)
    assert len(ids) == g["n_ids"] and sha(ids) == g["ids_sha256"], st
