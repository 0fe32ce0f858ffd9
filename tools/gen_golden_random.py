#!/usr/bin/env python3
"""Generate tests/golden/random_vectors.json: seeded random inputs and the outputs of the BIG-INTEGER oracle
(oracle/anemoi_ref.py, itself pinned to the reference's KATs) for every (field, instantiation):
permutation, compress (k = 2), compress_k (k = 4 on 4-3), hash_field of 5 elements, hash of 100 bytes,
merge, and a Merkle root over arity^2 leaves. Canonical integers as decimal strings. The GPU suite compares
the CUDA path with these committed vectors (no oracle code runs in that comparison)."""
import json
import os
import random
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import anemoi_ref as R

out = {}
for field in R.FIELDS:
    out[field] = {}
    for inst in R.INSTS:
        P = R.params(field, inst)
        rng = random.Random("golden-%s-%s" % (field, inst))
        W = P.width
        states = [[rng.randrange(P.p) for _ in range(W)] for _ in range(3)]
        states.append([P.p - 1] * W)          # boundary: all limbs at the top of the range
        rec = {"states": [[str(v) for v in s] for s in states], "permutation": [], "compress": [], "compress4": []}
        for s in states:
            t = list(s)
            R.permutation(P, t)
            rec["permutation"].append([str(v) for v in t])
            rec["compress"].append([str(v) for v in R.compress(P, s)])
            if W == 4:
                rec["compress4"].append([str(v) for v in R.compress_k(P, s, 4)])
        msg = [rng.randrange(P.p) for _ in range(5)]
        rec["hash_field_in"] = [str(v) for v in msg]
        rec["hash_field"] = str(R.hash_field(P, msg))
        data = bytes(rng.randrange(256) for _ in range(100))
        rec["hash_bytes_in"] = data.hex()
        rec["hash_bytes"] = str(R.hash_bytes(P, data))
        d = [rng.randrange(P.p) for _ in range(2)]
        rec["merge_in"] = [str(v) for v in d]
        rec["merge"] = str(R.merge(P, d[0], d[1]))
        leaves = [rng.randrange(P.p) for _ in range(W * W)]
        rec["merkle_leaves"] = [str(v) for v in leaves]
        rec["merkle_root"] = str(R.merkle_root(P, leaves, W))
        rec["digest_bytes"] = R.digest_to_bytes(P, d[0]).hex()
        out[field][inst] = rec
with open(os.path.join(ROOT, "tests", "golden", "random_vectors.json"), "w") as f:
    json.dump(out, f, indent=0, separators=(",", ":"))
    f.write("\n")
print("ok")
