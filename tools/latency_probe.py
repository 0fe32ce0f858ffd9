"""Developer probe: device time of ONE Jive compress launch over n states for small n (the sub-wave regime of the upper
Merkle levels and of small API calls). usage: python tools/latency_probe.py [field/inst substrings ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import anemoi_rust_b200 as A

only = sys.argv[1:] or ["pallas/anemoi_4_3", "bls12_377/anemoi_2_1", "bls12_381/anemoi_2_1", "bn_254/anemoi_4_3"]
dev = torch.device("cuda:0")
for (field, inst), H in sorted(A.HASHERS.items()):
    if not any(o in field + "/" + inst for o in only):
        continue
    f, W = H.FIELD, H.STATE_WIDTH
    host = f.random_mont((1 << 17) * W, seed=5)
    t_in = torch.from_numpy(host.view(np.int64)).to(dev)
    for lg in (0, 5, 8, 10, 11, 12, 13, 14, 15, 16, 17):
        n = 1 << lg
        x = t_in[: n * W]
        out = torch.empty((n, f.n64), dtype=torch.int64, device=dev)
        H.compress_k_batch(x, W, out=out)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            H.compress_k_batch(x, W, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"field": field, "inst": inst, "log2n": lg, "ms": round(best, 4), "us_per_state": round(best * 1e3 / n, 3)}), flush=True)
