import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import llmtokenizer_b200 as L
import parity_cases as pc
key = sys.argv[1]; P = int(sys.argv[2]); reps = int(sys.argv[3])
data, cap = pc.train_input(key)
exp = pc.expected_train(key)
om = exp["merges"]
for envs in ({},):
    for k in ("BPE_CUDA_PDL", "BPE_CUDA_SPECULATE", "BPE_CUDA_BATCH_MAX"):
        os.environ.pop(k, None)
    os.environ.update(envs)
    res = []
    for r in range(reps):
        try:
            m, t, st = L.train(data, max_merges=cap, n_gpus=P)
            k = min(len(m), len(om))
            bad = next((i for i in range(k) if tuple(m[i]) != tuple(om[i])), None)
            res.append("ok" if bad is None and len(m) == len(om) and pc.sha(t) == exp["ids_sha256"] else f"DIFF@{bad}")
        except Exception as e:
            res.append("ERR:" + str(e).split("flags")[-1][:6])
    print(key, P, envs, res, flush=True)
