"""GPU parity at the SHAPES of the BASELINE configs (not only their node functions): the 10 KB sponge messages of
config 2 (331 field elements at 31 bytes per element -- /root/reference benches/bn_254.rs:39-59 hashes 10 240
bytes; src/bn_254/anemoi_4_3/hasher.rs:19-129 is the code path), 10 240-byte strings through `hash` for a 31-byte
and a 47-byte field, fixed-length and ragged, and trees deep enough that most levels run as full-size launches
(arity 2 over 2^12 BLS12-377 leaves = config 4's node; arity 4 over 4^7 Pallas/Vesta leaves = config 3's node).
Everything is compared with the C oracle bit for bit, through the C ABI."""
import numpy as np
import pytest

import anemoi_rust_b200 as A
from anemoi_rust_b200 import HASHERS
from oracle import c_oracle as C

pytestmark = pytest.mark.gpu
SEED = 0xA7E301 + 2  # config index 2


def ids(field, inst):
    return A.FIELD_NAMES.index(field), (0 if inst == "anemoi_2_1" else 1)


@pytest.mark.parametrize("field,inst,n_msgs", [("bn_254", "anemoi_4_3", 96), ("bn_254", "anemoi_2_1", 48),
                                                ("bls12_381", "anemoi_4_3", 24), ("pallas", "anemoi_4_3", 64)])
def test_hash_field_331_felts(field, inst, n_msgs):
    """config 2's message: 331 elements -> 111 permutations on 4-3 (331 = 3*110 + 1: the padded last block), 331 on 2-1."""
    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f = H.FIELD
    L = 331
    x = f.random_mont(n_msgs * L, SEED)
    got = H.hash_field_batch(x, felts_per_msg=L)
    assert np.array_equal(got, C.hash_field(fi, ii, x, n_msgs, L))
    # 330 = 3 * 110: sigma = 1, NO padded block (anemoi_4_3/hasher.rs:100-127); 332: two elements in the last block
    for L2 in (330, 332):
        m = 8
        assert np.array_equal(H.hash_field_batch(x[: m * L2], felts_per_msg=L2), C.hash_field(fi, ii, x[: m * L2], m, L2))


@pytest.mark.parametrize("field,inst", [("bn_254", "anemoi_4_3"), ("bn_254", "anemoi_2_1")])
def test_hash_field_ragged_long(field, inst):
    """Ragged batch whose messages straddle the config-2 length (and include empty / 1-element ones): every lane of a
    warp runs a different number of permutations."""
    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f = H.FIELD
    rng = np.random.default_rng(SEED + 1)
    lens = np.concatenate([[331, 0, 1, 330, 332, 3, 111], rng.integers(0, 400, size=41)]).astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    x = f.random_mont(int(offs[-1]), SEED + 2)
    assert np.array_equal(H.hash_field_batch(x, offsets=offs), C.hash_field_ragged(fi, ii, x, offs))


@pytest.mark.parametrize("field,inst,n_msgs", [("bn_254", "anemoi_4_3", 48), ("bn_254", "anemoi_2_1", 16),
                                                ("bls12_381", "anemoi_4_3", 16), ("bls12_377", "anemoi_2_1", 8),
                                                ("vesta", "anemoi_4_3", 32)])
def test_hash_10kb_byte_strings(field, inst, n_msgs):
    """Sponge::hash on 10 240-byte strings (the reference's `hash 10KB` bench): 31-byte chunks -> 331 elements for the
    4-limb fields, 47-byte chunks -> 218 elements for the 6-limb fields; the last chunk is short (10 240 = 330*31 + 10
    = 217*47 + 41), so the 0x01 pad byte is exercised at full length."""
    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f = H.FIELD
    rng = np.random.default_rng(SEED + 3)
    nb = 10240
    data = rng.integers(0, 256, size=(n_msgs, nb), dtype=np.uint8)
    assert np.array_equal(H.hash_batch(data), C.hash_bytes(fi, ii, data, n_msgs, nb))
    # exact multiples of the chunk size (no pad byte) at the same scale
    nb2 = 330 * f.byte_chunk if f.byte_chunk == 31 else 217 * f.byte_chunk
    d2 = np.ascontiguousarray(data[:4, :nb2]) if nb2 <= nb else rng.integers(0, 256, size=(4, nb2), dtype=np.uint8)
    assert np.array_equal(H.hash_batch(d2), C.hash_bytes(fi, ii, d2, 4, nb2))
    # ragged around 10 KB in one launch
    lens = [10240, 10239, 10241, 0, 31, 47, 10230, 5000]
    msgs = [bytes(rng.integers(0, 256, size=n, dtype=np.uint8)) for n in lens]
    exp = np.concatenate([C.hash_bytes(fi, ii, np.frombuffer(m, dtype=np.uint8), 1, len(m)) for m in msgs])
    assert np.array_equal(H.hash_ragged(msgs), exp)


@pytest.mark.parametrize("field,inst,height", [("bls12_377", "anemoi_2_1", 12), ("bls12_381", "anemoi_2_1", 11),
                                                ("pallas", "anemoi_4_3", 7), ("vesta", "anemoi_4_3", 7),
                                                ("bn_254", "anemoi_4_3", 6)])
def test_deeper_trees_vs_oracle(field, inst, height):
    """config 3 / 4 node functions on trees of 2^12 / 4^7 leaves: root, retained tree and a few openings vs the oracle."""
    import torch

    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f, ar = H.FIELD, H.STATE_WIDTH
    n = ar ** height
    leaves = f.random_mont(n, SEED + 4)
    exp_root = C.merkle_root(fi, ii, ar, leaves)
    assert np.array_equal(H.merkle_root(leaves), exp_root)
    # device-resident sharded entry with a single rank (no communicator): same root
    from anemoi_rust_b200 import merkle

    t = torch.from_numpy(leaves.view(np.int64)).cuda()
    r = merkle.merkle_root_distributed(H, t)
    torch.cuda.synchronize()
    assert np.array_equal(r.cpu().numpy().view(np.uint64), exp_root)
    # openings of the first, last and a middle leaf verify against the root
    idx = np.array([0, n - 1, n // 3], dtype=np.uint64)
    root, paths = H.merkle_open(leaves, idx)
    assert np.array_equal(root, exp_root)
    roots = H.merkle_verify(leaves[idx.astype(np.int64)], idx, paths)
    assert np.array_equal(roots, np.repeat(exp_root, len(idx), axis=0))


@pytest.mark.parametrize("field,inst,height", [("pallas", "anemoi_4_3", 11), ("bls12_381", "anemoi_2_1", 21)])
def test_host_pointer_root_with_chunked_upload(field, inst, height):
    """Trees big enough (>= 2^20 level-1 nodes) that anemoi_b200_merkle_root uploads the leaves in chunks and hashes the
    first level of each chunk while the next one is still in flight: the root must equal the device-resident build of the
    same leaves (itself checked against the oracle on the smaller trees above), and a sub-tree root the oracle can afford."""
    import torch

    from anemoi_rust_b200 import merkle

    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f, ar = H.FIELD, H.STATE_WIDTH
    n = ar ** height
    leaves = f.random_mont(n, SEED + 5)
    got = H.merkle_root(leaves)
    dev = merkle.merkle_root_device(H, torch.from_numpy(leaves.view(np.int64)).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(got, dev.cpu().numpy().view(np.uint64))
    sub = ar ** (6 if ar == 4 else 11)
    assert np.array_equal(H.merkle_root(leaves[:sub]), C.merkle_root(fi, ii, ar, leaves[:sub]))
