#!/usr/bin/env python3
"""Benchmark of the BPE merge-loop hot path (BASELINE.json metric: BPE train merges/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--workload c1|c2|c3]

One "step" = one complete training run of the workload (widen, count, all merges) from the corpus.
  value : merges/s with the corpus already resident in HBM (CUDA events on the engine's stream,
          max over ranks)
  e2e   : the same through the C-ABI calls with HOST buffers: pinned corpus -> H2D, train, merges +
          ids -> D2H, all inside the timed region
Workloads:
  c2 : synthetic 100 MB Zipf-word ASCII corpus, 4,096 merges   (default at every N, BASELINE configs[1])
  c3 : synthetic 1 GB byte-level Zipf corpus, 32,000 merges, sharded over N GPUs with one NCCL
       all-reduce of the pair-count deltas per merge             (--workload c3, BASELINE configs[2])
  c1 : the reference's random_text.txt to exhaustion (parity configuration; L2-resident)
N>1 is launched by torch.distributed.run (one rank per GPU); torch.distributed is only the bootstrap
(NCCL id broadcast, barriers, max-over-ranks), the data path is the engine's own NCCL communicator.

--impl reference times the UNMODIFIED reference (oracle/_ref/ref_harness, compiled from the reference
sources) on the host cores, each step a bounded sample of the same workload (first merges).
"""
import argparse
import ctypes as C
import gzip
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": dict(desc="random_text.txt (1,048,576 B, bundled with the reference), merges to exhaustion", kind=None,
               size=1048576, seed=0, merges=0, ref_sample_merges=100),
    "c2": dict(desc="synthetic 100 MB Zipf-word ASCII corpus (zipf_ascii seed 1234), 4,096 merges", kind=0,
               size=100_000_000, seed=1234, merges=4096, ref_sample_merges=16),
    "c3": dict(desc="synthetic 1 GB byte-level Zipf corpus (zipf_bytes seed 4321), 32,000 merges", kind=1,
               size=1_000_000_000, seed=4321, merges=32000, ref_sample_merges=2),
}
REF_BIN = os.path.join(ROOT, "oracle/_ref/ref_harness")
ORACLE_CLI = os.path.join(ROOT, "oracle/_build/bpe_oracle_cli")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_corpus(w, pinned=True):
    """Corpus in (pinned) host memory as a numpy view."""
    import torch
    n = w["size"]
    t = torch.empty(n, dtype=torch.uint8, pin_memory=pinned and torch.cuda.is_available())
    arr = t.numpy()
    if w["kind"] is None:
        with gzip.open(os.path.join(ROOT, "tests/golden/random_text.txt.gz"), "rb") as f:
            arr[:] = np.frombuffer(f.read(), dtype=np.uint8)
    else:
        from llmtokenizer_b200 import _lib
        lib = _lib.load_corpus()
        assert lib.gen_corpus_fill(w["kind"], arr.ctypes.data, n, w["seed"], 50000 if w["kind"] == 0 else 65536) == 0
    return t, arr


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference_sample(corpus_path, merges):
    """One bounded sample of the reference's CPU path: the first `merges` merges of the workload."""
    out = f"/tmp/bench_ref_{os.getpid()}"
    if os.path.exists(REF_BIN):
        cmd, kind = [REF_BIN, corpus_path, str(merges), out], "reference"
    else:
        cmd, kind = [ORACLE_CLI, "train", corpus_path, str(merges), "1", out], "port"
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd}: rc={r.returncode} {r.stderr[-300:]}")
    f = r.stdout.split()
    done, secs = int(f[0]), float(f[2])
    for ext in (".merges", ".ids"):
        try:
            os.remove(out + ext)
        except OSError:
            pass
    return done, secs, kind


def dist_setup(n_gpus):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group(backend="gloo")
    assert world == n_gpus or world == 1, f"--gpus {n_gpus} but WORLD_SIZE={world}"
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def bench_reference(args, w, rank, world):
    if rank != 0:
        return
    _, arr = make_corpus(w, pinned=False)
    path = f"/tmp/bench_corpus_{args.workload}.bin"
    arr.tofile(path)
    sample = int(os.environ.get("BPE_BENCH_REF_MERGES") or w["ref_sample_merges"])
    for _ in range(args.warmup):
        run_reference_sample(path, sample)
    done_tot, sec_tot, kind = 0, 0.0, "reference"
    for _ in range(args.steps):
        done, secs, kind = run_reference_sample(path, sample)
        done_tot += done
        sec_tot += secs
    os.remove(path)
    val = done_tot / sec_tot
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "bpe_train_merges_per_sec", "value": val, "unit": "merges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec_tot / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "merges": w["merges"], "corpus_bytes": w["size"]},
        "cpu_baseline": {"value": val, "unit": "merges/s", "cores": cores, "kind": kind,
                         "threads": "16 workers + main, hard-coded (bpe.c:409)",
                         "sample": f"first {sample} merges of the same corpus per step (compress() timed with "
                                   f"clock_gettime, file read + widen included)"},
        "e2e": {"value": val, "unit": "merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_engine(args, w, rank, world, local):
    import torch
    import llmtokenizer_b200 as L
    from llmtokenizer_b200 import _lib
    lib = _lib.load()
    assert lib.bpe_cuda_device_count() > local, "no CUDA device: the engine has no CPU fallback"
    torch.cuda.set_device(local)
    pin_t, corpus = make_corpus(w)
    n = corpus.size
    lo, hi = n * rank // world, n * (rank + 1) // world
    shard = corpus[lo:hi]

    ctx = L.Context(local)
    if world > 1:
        import torch.distributed as dist
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt[:] = torch.frombuffer(bytearray(L.nccl_unique_id()), dtype=torch.uint8)
        dist.broadcast(idt, 0)
        ctx.set_comm(rank, world, bytes(idt.numpy().tobytes()))
    M = w["merges"]

    # ---- resident-in-HBM timing -----------------------------------------------------------------
    ctx.upload_ptr(shard.ctypes.data, shard.size)
    launches = 0
    for _ in range(args.warmup):
        st = ctx.train(M)
    sampler = ClockSampler(local)
    barrier(world)
    torch.cuda.synchronize()
    sampler.start()
    dev_ms, merges_done = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = ctx.train(M)
        dev_ms += st["ms_device"]
        merges_done += st["n_merges"]
        launches += st["kernel_launches"]
    torch.cuda.synchronize()
    barrier(world)
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    dev_ms = max_over_ranks(dev_ms, world)
    wall_ms = max_over_ranks(wall_ms, world)
    value = merges_done / (dev_ms * 1e-3)
    final_tokens = st["n_tokens"]

    # ---- end to end through host buffers ----------------------------------------------------------
    nm, nt = ctx.result_sizes()
    out_m = torch.empty((max(nm, 1), 2), dtype=torch.int32, pin_memory=True)
    out_t = torch.empty(max(nt + 1024, 1), dtype=torch.int32, pin_memory=True)

    def e2e_step():
        ctx.upload_ptr(shard.ctypes.data, shard.size)        # H2D of this step's input
        s = ctx.train(M)
        ctx.download_into(out_m.data_ptr(), out_t.data_ptr())  # D2H of merges + ids
        return s

    e2e_step()
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_merges = 0
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        s = e2e_step()
        e2e_merges += s["n_merges"]
        launches += s["kernel_launches"]
    torch.cuda.synchronize()
    barrier(world)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world)
    e2e_value = e2e_merges / (e2e_ms * 1e-3)
    h2d = int(shard.size)
    d2h = int(nm * 8 + nt * 4)

    # ---- roofline of the dominant kernel (replace+scan+delta), one extra profiled step -----------
    ctx.set_option("profile_replace", 1)
    sp = ctx.train(M)
    ctx.set_option("profile_replace", 0)
    peak, peak_src = peaks()
    k_ms = sp["replace_ms"]
    k_bytes = sp["replace_bytes"] / world  # algorithmic bytes on this rank ~ global / N
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": "replace_stream_kernel (fused replace + prefix-scan compaction + pair-count deltas; a == b passes: replace_kernel)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "traffic": None, "launches": sp["replace_passes"], "avg_launch_us": 1e3 * k_ms / max(1, sp["replace_passes"]),
        "merges_per_pass": sp["n_merges"] / max(1, sp["replace_passes"]),
        "algorithmic_bytes_per_step": sp["replace_bytes"],
        "kernel_share_of_step": k_ms / sp["ms_device"] if sp["ms_device"] else None,
        "other_kernels_ms": {"apply_select": sp["apply_ms"] + sp["select_ms"], "host_gaps": sp["gap_ms"]},
        "how": "CUDA events around every pass (replace_stream_kernel / replace_kernel launch) in one extra step of the same "
               "workload (events off in the timed steps); bytes = sum over PASSES of 4*(tokens in + tokens out) - a pass "
               "that carries several provably-next merges is counted once; traffic: see profiles/README.md (ncu: DRAM "
               "bytes = algorithmic bytes, no re-reads)",
    }
    whole = (sp["replace_bytes"] / world + 9 * shard.size) / (sp["ms_device"] * 1e-3) / 1e9
    roofline["whole_step_gbs"] = whole
    roofline["whole_step_frac"] = whole / peak

    # ---- encode GB/s: apply the learned merge table to the same corpus (resident; N = 1 only) ------
    encode = None
    if world == 1:
        merges_np, _ = ctx.download(tokens=False)
        se = ctx.encode(merges_np)
        se = ctx.encode(merges_np)
        encode = {"value": shard.size / (se["ms_device"] * 1e-3) / 1e9, "unit": "GB/s of input", "ranks": int(len(merges_np)),
                  "ms": se["ms_device"], "passes": se["replace_passes"]}

    # ---- full-size self-checks (N = 1, outside every timed region): size-independent properties ----
    checks = None
    if world == 1:
        import hashlib
        ctx.train(M)
        m_b, t_b = ctx.download()
        nbytes = ctx.decode(m_b, download=False)       # ids -> bytes on the device, compared with the shard there
        mism = ctx.decode_mismatches() if nbytes == shard.size else -1
        ctx.encode(m_b)
        _, t_e = ctx.download(merges=False)
        ctx.set_option("batch_max", 1)                 # one merge per pass: the sequential order by construction
        ctx.train(M)
        m_1, t_1 = ctx.download()
        ctx.set_option("batch_max", 8)
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        checks = {"decode_round_trip_mismatching_bytes": int(mism),
                  "batched_passes_equal_one_merge_per_pass": bool(sha(m_b) == sha(m_1) and sha(t_b) == sha(t_1)),
                  "encode_with_learned_merges_reproduces_training_ids": bool(sha(t_e) == sha(t_b)),
                  "merges_sha256": sha(m_b)[:16], "ids_sha256": sha(t_b)[:16]}
        gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "c2_full.json")
        if args.workload == "c2" and os.path.exists(gold):
            g = json.load(open(gold))       # digests of the CPU oracle's result for this very corpus (tools/make_c2_golden.py)
            checks["equals_oracle_digest"] = bool(sha(m_b) == g["merges_sha256"] and sha(t_b) == g["ids_sha256"])

    # ---- CPU reference beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        path = f"/tmp/bench_corpus_{args.workload}.bin"
        corpus.tofile(path)
        done, secs, kind = run_reference_sample(path, w["ref_sample_merges"])
        os.remove(path)
        cpu = {"value": done / secs, "unit": "merges/s", "cores": os.cpu_count(), "kind": kind,
               "threads": "16 workers + main, hard-coded (bpe.c:409)",
               "sample": f"first {done} merges of the same corpus ({secs:.1f} s, compress() incl. file read + widen)"}

    if rank == 0:
        line = {
            "metric": "bpe_train_merges_per_sec", "value": value, "unit": "merges/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "merges": M, "corpus_bytes": n,
                       "shards": world, "final_tokens": int(final_tokens),
                       "l2": "every step starts from the resident byte corpus and re-widens it into a 4 B/token stream (400 MB for c2, larger than the 126 MB L2); passes compact the stream in place, so late passes (76-100 MB) reuse what the previous pass left in L2 - that reuse is part of the algorithm, nothing survives from one timed step to the next; no flush"
                       if n >= 50_000_000 else "stream fits in L2 (parity configuration, not a bandwidth one)",
                       "wall_ms_per_step": wall_ms / args.steps},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "merges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "encode": encode,
            "checks": checks,
            "engine_stats": {k: sp[k] for k in ("same_bucket_ties", "threshold_edges", "resolver_runs", "census_runs",
                                                 "table_rehashes", "table_capacity", "final_distinct", "replace_passes",
                                                 "batch_merges", "batch_passes")},
        }
        print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload is None:
        # BASELINE.json quotes the metric on configs[1] (c2); every N runs it so that the per-N values are
        # comparable.  The 1 GB scaling configuration is `--workload c3` (numbers in DESIGN.md / profiles/).
        args.workload = os.environ.get("BPE_BENCH_WORKLOAD") or "c2"
    w = WORKLOADS[args.workload]
    rank, world, local = dist_setup(args.gpus)
    if args.impl == "reference":
        bench_reference(args, w, rank, world)
    else:
        bench_engine(args, w, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
