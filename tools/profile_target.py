"""ncu target: a few launches of one (field, instantiation) Jive compress on device-resident data.
usage: python tools/profile_target.py <field> <2_1|4_3> <log2n> [launches]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import anemoi_rust_b200 as A

field, inst, log2n = sys.argv[1], "anemoi_" + sys.argv[2], int(sys.argv[3])
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
H = A.HASHERS[(field, inst)]
f, W = H.FIELD, H.STATE_WIDTH
n = 1 << log2n
x = torch.from_numpy(f.random_mont(n * W, 0xA7E301).view(np.int64)).cuda()
out = torch.empty((n, f.n64), dtype=torch.int64, device="cuda")
for _ in range(launches):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    H.compress_k_batch(x, W, out=out)
    e1.record()
    torch.cuda.synchronize()
    print("%s %s n=2^%d: %.3f ms, %.3f M/s" % (field, inst, log2n, e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e3))
