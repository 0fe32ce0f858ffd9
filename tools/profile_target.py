"""ncu target: a few launches of one (field, instantiation) kernel on device-resident data.
usage: python tools/profile_target.py <field> <2_1|4_3> <log2n> [launches] [mode]
  mode = compress (default; Jive k = STATE_WIDTH on 2^log2n states) | hash<L> (sponge hash_field, 2^log2n messages of L elements)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import anemoi_rust_b200 as A

field, inst, log2n = sys.argv[1], "anemoi_" + sys.argv[2], int(sys.argv[3])
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mode = sys.argv[5] if len(sys.argv) > 5 else "compress"
H = A.HASHERS[(field, inst)]
f, W = H.FIELD, H.STATE_WIDTH
n = 1 << log2n
per = int(mode[4:]) if mode.startswith("hash") else W
g = torch.Generator(device="cuda")
g.manual_seed(0xA7E301)
x = torch.randint(-(1 << 63), (1 << 63) - 1, (n * per, f.n64), dtype=torch.int64, device="cuda", generator=g)
x[:, f.n64 - 1] &= (1 << (f.p.bit_length() - 1 - 64 * (f.n64 - 1))) - 1
out = torch.empty((n, f.n64), dtype=torch.int64, device="cuda")
for _ in range(launches):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if mode.startswith("hash"):
        H.hash_field_batch(x, felts_per_msg=per)
    else:
        H.compress_k_batch(x, W, out=out)
    e1.record()
    torch.cuda.synchronize()
    print("%s %s %s n=2^%d: %.3f ms, %.3f M units/s" % (field, inst, mode, log2n, e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e3))
