"""Phase timers of apply_select_kernel (device globaltimer, BPE_CUDA_DEBUG) for a workload: python tools/phase_timers.py c2|c3 [merges]"""
import os, sys
PLAIN = "--plain" in sys.argv      # one training run, no debug output: the command ncu wraps
if PLAIN:
    sys.argv.remove("--plain")
else:
    os.environ["BPE_CUDA_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import llmtokenizer_b200 as L
w = bench.WORKLOADS[sys.argv[1]]
M = int(sys.argv[2]) if len(sys.argv) > 2 else w["merges"]
arr = np.empty(w["size"], dtype=np.uint8)
bench.fill_corpus(w, arr, 0, w["size"])
ctx = L.Context(0)
ctx.upload(arr)
if not PLAIN:
    ctx.train(M)
st = ctx.train(M)
print({k: st[k] for k in ("n_merges", "replace_passes", "ms_device", "table_capacity", "final_distinct")})
if not PLAIN:
    ctx.set_option("profile_replace", 1)
    sp = ctx.train(M)
    print("profiled step:", {k: round(sp[k], 1) for k in ("ms_device", "replace_ms", "apply_ms", "select_ms", "gap_ms")})
