"""Where the one-call entry points spend their time: python tools/onecall_timing.py c2|c3"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import llmtokenizer_b200 as L
w = bench.WORKLOADS[sys.argv[1]]
arr = np.empty(w["size"], dtype=np.uint8)
bench.fill_corpus(w, arr, 0, w["size"])
path = "/tmp/onecall.bin"
arr.tofile(path)
for rep in range(2):
    t0 = time.perf_counter(); m, t, st = L.train(arr, max_merges=w["merges"]); dt = time.perf_counter() - t0
    print("train(memory)", round(dt, 3), {k: round(st[k], 1) for k in ("ms_h2d", "ms_device", "ms_d2h", "ms_total")}, flush=True)
    t0 = time.perf_counter(); m, t, st = L.train_file(path, max_merges=w["merges"]); dt = time.perf_counter() - t0
    print("train_file   ", round(dt, 3), {k: round(st[k], 1) for k in ("ms_h2d", "ms_device", "ms_d2h", "ms_total")}, flush=True)
