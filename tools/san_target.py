"""compute-sanitizer target: every kernel mode once on small, odd-sized batches (tail lanes, ragged, bytes,
Merkle open/verify) for one 12-limb and one 8-limb field."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

import anemoi_rust_b200 as A

for name in ("AnemoiBls12_381_2_1", "AnemoiBls12_381_4_3", "AnemoiPallas_2_1", "AnemoiPallas_4_3"):
    H = getattr(A, name)
    f, W = H.FIELD, H.STATE_WIDTH
    x = f.random_mont(37 * W, 1)
    H.permutation_batch(x)
    H.sbox_layer_batch(x)
    H.compress_batch(x)
    if W == 4:
        H.compress_k_batch(x, 4)
    H.hash_field_batch(f.random_mont(11 * 5, 2), felts_per_msg=5)
    offs = np.array([0, 0, 1, 3, 6, 10, 17], dtype=np.uint64)
    H.hash_field_batch(f.random_mont(17, 3), offsets=offs)
    H.hash_batch(np.random.default_rng(4).integers(0, 256, size=(7, 101), dtype=np.uint8))
    H.merge_batch(f.random_mont(2 * 9, 5))
    H.Digest([5]).to_bytes()
    leaves = f.random_mont(W ** 3, 6)
    H.merkle_root(leaves)
    root, paths = H.merkle_open(leaves, [0, 5, W ** 3 - 1])
    H.merkle_verify(leaves[[0, 5, W ** 3 - 1]], [0, 5, W ** 3 - 1], paths)
    print(name, "ok", flush=True)
