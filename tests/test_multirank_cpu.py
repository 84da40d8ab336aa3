"""The N > 1 host-side plumbing on CPU (gloo, world size 2): rank bootstrap as bench.py does it (shard
bounds, broadcast of the 128-byte communicator id, max-over-ranks timing), the reference arm's
"rank 0 alone runs and prints" contract, and the shard-edge protocol itself: two ranks that only
exchange the 8-word edge records must reproduce the single-stream rewrite."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["BPE_ROOT"]); sys.path.insert(0, os.path.join(os.environ["BPE_ROOT"], "tests"))
import bench, oracle_api
rank, world, local = bench.dist_setup(2)
assert world == 2
# the bootstrap bench_engine() performs
idt = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idt[:] = torch.arange(128, dtype=torch.uint8)
dist.broadcast(idt, 0)
assert idt.tolist() == list(range(128))
assert bench.max_over_ranks(1.0 + rank, world) == 2.0
assert bench.sum_over_ranks(1.0 + rank, world) == 3.0
bench.barrier(world)
# the digest of the ranks' id arrays in rank order (how bench.py checks sharded results against the oracle's digest)
import hashlib
parts = [np.arange(5, dtype=np.uint32), np.arange(100, 107, dtype=np.uint32)]
got = bench.sha_over_ranks(parts[rank], rank, world)
if rank == 0:
    assert got == hashlib.sha256(np.concatenate(parts).astype("<u4").tobytes()).hexdigest()
# every rank generates only its own shard of a chunked corpus (configs 4 and 5): together they are the corpus
w = dict(kind=1, size=10_500, seed=3, chunk=1000)
whole = np.empty(w["size"], dtype=np.uint8); bench.fill_corpus(w, whole, 0, w["size"])
lo, hi = w["size"] * rank // world, w["size"] * (rank + 1) // world
mine = np.empty(hi - lo, dtype=np.uint8); bench.fill_corpus(w, mine, lo, hi)
assert np.array_equal(mine, whole[lo:hi]) and whole.min() > 0
# shard bounds + edge records: each rank rewrites its shard with the tokens it gets from the other rank's record
oracle = oracle_api.load()
rng = np.random.default_rng(5)
stream = rng.integers(97, 100, 20001, dtype=np.uint32)
a, b, z = 97, 98, 256
n = stream.size
lo, hi = n * rank // world, n * (rank + 1) // world
mine = stream[lo:hi]
rec = torch.zeros(world, 8, dtype=torch.int64)   # len, first three, last two (same fields as the GPU records)
rec[rank, 0] = mine.size
rec[rank, 1:4] = torch.from_numpy(mine[:3].astype(np.int64))
rec[rank, 4:6] = torch.from_numpy(mine[-2:].astype(np.int64))
dist.all_reduce(rec)                               # disjoint slots: the sum is an all-gather
ext = list(mine)
drop_first = False
if rank > 0:                                       # did the left neighbour's last token start a replacement with my first?
    drop_first = int(rec[rank - 1, 5]) == a and int(mine[0]) == b
if rank < world - 1:
    ext = ext + [int(rec[rank + 1, 1])]            # one token of look-ahead decides my last position
out = oracle.rewrite(np.array(ext, dtype=np.uint32), a, b, z)
if rank < world - 1:
    # the appended halo token is not mine: unless my last token started a replacement with it (then the output
    # ends in z, which is mine), it is still the last output token and goes away again
    if not (int(mine[-1]) == a and ext[-1] == b):
        out = out[:-1]
if drop_first:
    out = out[1:]
sizes = torch.zeros(world, dtype=torch.int64); sizes[rank] = out.size; dist.all_reduce(sizes)
full = torch.zeros(int(sizes.sum()), dtype=torch.int64)
off = int(sizes[:rank].sum())
full[off:off + out.size] = torch.from_numpy(out.astype(np.int64))
dist.all_reduce(full)
if rank == 0:
    want = oracle.rewrite(stream, a, b, z)
    assert np.array_equal(full.numpy().astype(np.uint32), want), "sharded rewrite differs from the single stream"
    print("MULTIRANK_OK")
dist.destroy_process_group()
"""


def run_two_ranks(script_args, env_extra=None, timeout=300):
    env = dict(os.environ, BPE_ROOT=ROOT, **(env_extra or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611"] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=timeout, cwd=ROOT)


def test_two_rank_bootstrap_and_edge_protocol(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = run_two_ranks([str(script)])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MULTIRANK_OK" in r.stdout


def test_reference_arm_runs_on_rank0_only():
    if not os.path.exists(os.path.join(ROOT, "oracle/_ref/ref_harness")) and not os.path.exists(
            os.path.join(ROOT, "oracle/_build/bpe_oracle_cli")):
        pytest.skip("neither the compiled reference nor the oracle CLI is built")
    r = run_two_ranks(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "c1"],
                      env_extra={"BPE_BENCH_REF_MERGES": "3"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line (rank 0)"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
