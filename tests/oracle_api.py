"""TEST INFRASTRUCTURE: ctypes view of the CPU oracle (oracle/bpe_oracle.c) and golden-fixture helpers.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes as C
import gzip
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle/_build/libbpe_oracle.so")
GOLD = os.path.join(ROOT, "tests/golden")

FAITHFUL, FAST, FAST_CF = 0, 1, 2


class OStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_merges", "n_tokens", "same_bucket_ties", "threshold_edges", "faithful_iters",
                                          "census_iters", "final_distinct")] + [("thread_buckets", C.c_uint64 * 16)]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        P = C.POINTER
        lib.bo_train.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, P(C.c_void_p), P(C.c_size_t), P(C.c_void_p),
                                 P(C.c_size_t), P(OStats)]
        lib.bo_train.restype = C.c_int
        lib.bo_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, P(C.c_void_p), P(C.c_size_t)]
        lib.bo_encode.restype = C.c_int
        lib.bo_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, P(C.c_void_p), P(C.c_size_t)]
        lib.bo_decode.restype = C.c_int
        lib.bo_murmur3_pair.argtypes = [C.c_uint32, C.c_uint32]
        lib.bo_murmur3_pair.restype = C.c_uint32
        lib.bo_merged_buckets.argtypes = [C.c_uint64]
        lib.bo_merged_buckets.restype = C.c_uint64
        lib.bo_rewrite.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.bo_rewrite.restype = C.c_size_t
        lib.bo_rewrite_sharded.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_size_t,
                                           C.c_void_p, C.c_void_p]
        lib.bo_rewrite_sharded.restype = C.c_size_t
        lib.bo_free.argtypes = [C.c_void_p]
        lib.bo_set_workers.argtypes = [C.c_int, C.c_size_t]
        lib.bo_set_workers.restype = None
        lib.bo_set_candidate_floor.argtypes = [C.c_uint32]
        lib.bo_set_candidate_floor.restype = None

    def configure(self, workers=1, min_tokens=1 << 22, candidate_floor=64):
        """helper threads / candidate-list shortcut of the FAST modes (the defaults are the library's)"""
        self.lib.bo_set_workers(workers, min_tokens)
        self.lib.bo_set_candidate_floor(candidate_floor)

    def train(self, data, max_merges=0, mode=FAST_CF):
        arr = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
        mp, tp = C.c_void_p(), C.c_void_p()
        nm, nt = C.c_size_t(), C.c_size_t()
        st = OStats()
        rc = self.lib.bo_train(arr.ctypes.data if arr.size else None or C.c_char_p(b""), arr.size, max_merges, mode,
                               C.byref(mp), C.byref(nm), C.byref(tp), C.byref(nt), C.byref(st))
        if rc:
            return rc, None, None, None
        m = np.ctypeslib.as_array(C.cast(mp, C.POINTER(C.c_uint32)), shape=(nm.value, 2)).copy() if nm.value else np.zeros((0, 2), np.uint32)
        t = np.ctypeslib.as_array(C.cast(tp, C.POINTER(C.c_uint32)), shape=(nt.value,)).copy() if nt.value else np.zeros(0, np.uint32)
        self.lib.bo_free(mp)
        self.lib.bo_free(tp)
        stats = {n: getattr(st, n) for n, _ in OStats._fields_[:-1]}
        stats["thread_buckets"] = list(st.thread_buckets)
        return 0, m, t, stats

    def encode(self, data, merges):
        arr = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
        mg = np.ascontiguousarray(np.asarray(merges, dtype=np.uint32).reshape(-1, 2))
        tp, nt = C.c_void_p(), C.c_size_t()
        rc = self.lib.bo_encode(arr.ctypes.data if arr.size else C.c_char_p(b""), arr.size, mg.ctypes.data if mg.size else None,
                                mg.shape[0], C.byref(tp), C.byref(nt))
        assert rc == 0, rc
        t = np.ctypeslib.as_array(C.cast(tp, C.POINTER(C.c_uint32)), shape=(nt.value,)).copy() if nt.value else np.zeros(0, np.uint32)
        self.lib.bo_free(tp)
        return t

    def decode(self, ids, merges):
        ids = np.ascontiguousarray(np.asarray(ids, dtype=np.uint32))
        mg = np.ascontiguousarray(np.asarray(merges, dtype=np.uint32).reshape(-1, 2))
        bp, nb = C.c_void_p(), C.c_size_t()
        rc = self.lib.bo_decode(ids.ctypes.data if ids.size else None, ids.size, mg.ctypes.data if mg.size else None, mg.shape[0],
                                C.byref(bp), C.byref(nb))
        assert rc == 0, rc
        out = C.string_at(bp, nb.value)
        self.lib.bo_free(bp)
        return out

    def rewrite(self, toks, a, b, z):
        toks = np.ascontiguousarray(toks, dtype=np.uint32)
        out = np.zeros(max(1, toks.size), dtype=np.uint32)
        m = self.lib.bo_rewrite(toks.ctypes.data, toks.size, a, b, z, out.ctypes.data)
        return out[:m].copy()

    def rewrite_sharded(self, toks, a, b, z, shards, tile):
        toks = np.ascontiguousarray(toks, dtype=np.uint32)
        out = np.zeros(max(1, toks.size), dtype=np.uint32)
        delta = np.zeros(4 * (z + 1), dtype=np.int32)
        m = self.lib.bo_rewrite_sharded(toks.ctypes.data, toks.size, a, b, z, shards, tile, out.ctypes.data, delta.ctypes.data)
        return out[:m].copy(), delta.reshape(-1, 4)


def load():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return Oracle(C.CDLL(ORACLE_SO))


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLD) if f.endswith(".npz"))


def golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    g = {"cap": int(z["cap"]), "status": int(z["status"]), "merges": z["merges"].astype(np.uint32)}
    if "input_ref" in z.files:
        with gzip.open(os.path.join(GOLD, str(z["input_ref"])), "rb") as f:
            g["input"] = np.frombuffer(f.read(), dtype=np.uint8)
        g["ids"] = None
        g["n_ids"] = int(z["n_ids"])
        g["ids_sha256"] = str(z["ids_sha256"])
    else:
        g["input"] = z["input"].astype(np.uint8)
        g["ids"] = z["ids"].astype(np.uint32)
        g["n_ids"] = len(g["ids"])
        g["ids_sha256"] = hashlib.sha256(g["ids"].astype("<u4").tobytes()).hexdigest()
    return g


def ids_sha(ids):
    return hashlib.sha256(np.asarray(ids, dtype="<u4").tobytes()).hexdigest()


def pair_counts(toks):
    """Overlapping-occurrence pair histogram of a token stream (what bpe.c:460-471 counts)."""
    toks = np.asarray(toks, dtype=np.uint64)
    if toks.size < 2:
        return {}
    keys = toks[:-1] | (toks[1:] << np.uint64(32))
    u, c = np.unique(keys, return_counts=True)
    return dict(zip(u.tolist(), c.tolist()))
