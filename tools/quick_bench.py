"""Developer timing sweep: device-resident Jive compress for every (field, instantiation). Not the
contract benchmark (that is bench.py) -- used to steer kernel work."""
import argparse
import ctypes
import json
import sys
import os

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import anemoi_rust_b200 as A
from anemoi_rust_b200 import ffi

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
n = 1 << args.log2n
dev = torch.device("cuda:0")
for (field, inst), H in sorted(A.HASHERS.items()):
    if args.only and args.only not in field + "/" + inst:
        continue
    f = H.FIELD
    W = H.STATE_WIDTH
    k = W  # 2-1: k=2, 4-3: k=4
    host = f.random_mont(n * W, seed=0xA7E301)
    t_in = torch.from_numpy(host.view(np.int64)).to(dev)
    t_out = torch.empty((n, f.n64), dtype=torch.int64, device=dev)
    H.compress_k_batch(t_in, k, out=t_out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        H.compress_k_batch(t_in, k, out=t_out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"field": field, "inst": inst, "n": n, "ms": round(best, 3), "Mcompress_per_s": round(n / best / 1e3, 4)}), flush=True)
