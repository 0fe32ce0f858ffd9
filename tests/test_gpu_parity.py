"""GPU parity, part 2: the CUDA path (through the C ABI) against the C oracle on seeded random inputs, for
every (field, instantiation) and every entry point, including ragged/empty inputs, tails that do not fill
a block, device-pointer calls on torch tensors, the Merkle builder, and error behaviour. Bit-exact:
all comparisons are on Montgomery limbs."""
import ctypes

import numpy as np
import pytest

import anemoi_rust_b200 as A
from anemoi_rust_b200 import HASHERS, ffi
from oracle import c_oracle as C

pytestmark = pytest.mark.gpu

CASES = sorted(HASHERS)
SEED = 0xA7E301


def ids(field, inst):
    return A.FIELD_NAMES.index(field), (0 if inst == "anemoi_2_1" else 1)


@pytest.mark.parametrize("field,inst", CASES)
def test_permute_compress_random(field, inst):
    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f, W = H.FIELD, H.STATE_WIDTH
    n = 4099  # not a multiple of the block size: exercises the shadowed tail lanes
    x = f.random_mont(n * W, SEED + 1)
    assert np.array_equal(H.permutation_batch(x), C.permute(fi, ii, x))
    assert np.array_equal(H.compress_batch(x), C.compress(fi, ii, 2, x))
    if W == 4:
        assert np.array_equal(H.compress_k_batch(x, 4), C.compress(fi, ii, 4, x))
    # small batches take the one-warp-per-block path
    for m in (1, 2, 31, 33):
        assert np.array_equal(H.compress_batch(x[: m * W]), C.compress(fi, ii, 2, x[: m * W]))


@pytest.mark.parametrize("field,inst", CASES)
def test_sponge_random(field, inst):
    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f = H.FIELD
    rng = np.random.default_rng(SEED + 2)
    for L in (0, 1, 2, 3, 5, 6, 7):
        n = 70
        x = f.random_mont(max(n * L, 1), SEED + 3 + L)[: n * L]
        got = H.hash_field_batch(x, felts_per_msg=L, n_msgs=n)
        exp = C.hash_field(fi, ii, x, n, L)
        assert got.shape == exp.shape and np.array_equal(got, exp)
        if L == 0:
            assert not got.any()   # hash_field([]) == 0 (SURVEY Q3 / sigma-only state)
    # ragged
    lens = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 13, 1, 3], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    x = f.random_mont(int(offs[-1]), SEED + 20)
    assert np.array_equal(H.hash_field_batch(x, offsets=offs), C.hash_field_ragged(fi, ii, x, offs))
    # ragged with a non-zero first offset
    offs2 = offs[3:]
    assert np.array_equal(H.hash_field_batch(x, offsets=offs2), C.hash_field_ragged(fi, ii, x, offs2))
    # bytes: lengths around the 31/47-byte chunk boundaries
    B = f.byte_chunk
    for nb in (0, 1, B - 1, B, B + 1, 2 * B, 3 * B - 1, 3 * B, 3 * B + 1, 200):
        n = 9
        data = rng.integers(0, 256, size=(n, nb), dtype=np.uint8)
        got = H.hash_batch(data, bytes_per_msg=nb) if nb else H.hash_batch(np.zeros((n, 0), dtype=np.uint8))
        assert np.array_equal(got, C.hash_bytes(fi, ii, data, n, nb))
    # ragged byte strings in one launch
    msgs = [bytes(rng.integers(0, 256, size=int(L), dtype=np.uint8)) for L in (0, 1, B, B + 1, 5 * B, 7, 3 * B - 1, 0, 1000)]
    exp = np.concatenate([C.hash_bytes(fi, ii, np.frombuffer(m, dtype=np.uint8), 1, len(m)) for m in msgs])
    assert np.array_equal(H.hash_ragged(msgs), exp)
    # merge + digest bytes
    d = f.random_mont(2 * 37, SEED + 30)
    assert np.array_equal(H.merge_batch(d), C.merge(fi, ii, d))
    out = np.empty(37 * f.felt_bytes, dtype=np.uint8)
    ffi.check(ffi.lib.anemoi_b200_digest_to_bytes(f.id, ctypes.c_void_p(d.ctypes.data), ctypes.c_void_p(out.ctypes.data), 37, 0))
    assert np.array_equal(out, C.digest_to_bytes(fi, d[:37]))


@pytest.mark.parametrize("field,inst", CASES)
def test_merkle(field, inst):
    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f, ar = H.FIELD, H.STATE_WIDTH
    for h in (0, 1, 2, 5 if ar == 4 else 9):
        n = ar ** h
        leaves = f.random_mont(n, SEED + 40 + h)
        assert np.array_equal(H.merkle_root(leaves), C.merkle_root(fi, ii, ar, leaves))
    with pytest.raises(A.LengthError):
        H.merkle_root(f.random_mont(ar * 3, 1))


def test_merkle_reduce_dev_partial_levels():
    import torch
    from anemoi_rust_b200.merkle import merkle_reduce

    H = A.AnemoiPallas_4_3
    fi, ii = ids("pallas", "anemoi_4_3")
    f = H.FIELD
    n = 2 * 4 ** 4  # two sub-trees: what one rank of a 2-GPU run holds for a 4^5-leaf tree
    leaves = f.random_mont(n, SEED + 50)
    t = torch.from_numpy(leaves.view(np.int64)).cuda()
    part = merkle_reduce(H, t, levels=4).cpu().numpy().view(np.uint64)
    exp = np.concatenate([C.merkle_root(fi, ii, 4, leaves[: n // 2]), C.merkle_root(fi, ii, 4, leaves[n // 2:])])
    assert np.array_equal(part, exp)
    with pytest.raises(A.LengthError):
        merkle_reduce(H, t, levels=5)


@pytest.mark.parametrize("field,inst", [("bls12_381", "anemoi_2_1"), ("vesta", "anemoi_4_3")])
def test_device_pointer_path_matches_host_path(field, inst):
    import torch

    H = HASHERS[(field, inst)]
    f, W = H.FIELD, H.STATE_WIDTH
    n = 1000
    x = f.random_mont(n * W, SEED + 60)
    t = torch.from_numpy(x.view(np.int64)).cuda()
    out = H.compress_k_batch(t, W)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), H.compress_k_batch(x, W))
    # unaligned (8-byte but not 16-byte aligned) device views take the 64-bit load path
    pad = torch.zeros(n * W * f.n64 + 1, dtype=torch.int64, device="cuda")
    pad[1:] = t.reshape(-1)
    view = pad[1:].reshape(n * W, f.n64)
    assert view.data_ptr() % 16 == 8
    out2 = H.compress_k_batch(view, W)
    assert torch.equal(out2, out)
    p = H.permutation_batch(t.clone())
    assert np.array_equal(p.cpu().numpy().view(np.uint64), H.permutation_batch(x))
    hf = H.hash_field_batch(t.reshape(-1, 5, f.n64))
    assert np.array_equal(hf.cpu().numpy().view(np.uint64), H.hash_field_batch(x.reshape(-1, 5, f.n64)))


def test_error_behaviour_matches_reference_asserts():
    H2, H4 = A.AnemoiBls12_381_2_1, A.AnemoiBls12_381_4_3
    with pytest.raises(AssertionError):   # hasher.rs:97 assert!(elems.len() == STATE_WIDTH)
        H2.compress([1, 2, 3])
    with pytest.raises(AssertionError):   # hasher.rs:107 assert!(k == 2)
        H2.compress_k([1, 2], 4)
    with pytest.raises(AssertionError):   # 4-3 hasher.rs:163-165
        H4.compress_k([1, 2, 3, 4], 3)
    with pytest.raises(AssertionError):
        H4.compress_k([1, 2, 3, 4], 1)
    lib = ffi.lib
    z = np.zeros(64, dtype=np.uint64)
    p = ctypes.c_void_p(z.ctypes.data)
    assert lib.anemoi_b200_compress(9, 0, 2, p, p, 1, 0) == ffi.ERR_FIELD
    assert lib.anemoi_b200_compress(1, 2, 2, p, p, 1, 0) == ffi.ERR_INST
    assert lib.anemoi_b200_compress(1, 0, 4, p, p, 1, 0) == ffi.ERR_ARITY
    assert lib.anemoi_b200_compress(1, 1, 3, p, p, 1, 0) == ffi.ERR_ARITY
    assert lib.anemoi_b200_compress(1, 0, 2, None, p, 1, 0) == ffi.ERR_ARG
    assert lib.anemoi_b200_compress(1, 0, 2, p, p, 1, 99) == ffi.ERR_ARG
    assert lib.anemoi_b200_compress(1, 0, 2, p, p, 0, 0) == ffi.OK  # empty batch
    assert lib.anemoi_b200_merkle_root(1, 0, 4, p, 4, p, 1) == ffi.ERR_ARITY
    assert lib.anemoi_b200_merkle_root(1, 0, 2, p, 6, p, 1) == ffi.ERR_LENGTH
    assert lib.anemoi_b200_merkle_root(1, 0, 2, p, 0, p, 1) == ffi.ERR_LENGTH


def test_full_size_config1_properties():
    """BASELINE config 1 at full size (BLS12-381 Anemoi-2-1, 2^20 pairs): sampled oracle check,
    determinism and position independence."""
    H = A.AnemoiBls12_381_2_1
    f = H.FIELD
    n = 1 << 20
    x = f.random_mont(2 * n, SEED)
    out = H.compress_batch(x)
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(n, size=2048, replace=False))
    sample = x.reshape(n, 2, f.n64)[idx].reshape(-1, f.n64)
    assert np.array_equal(out[idx], C.compress(1, 0, 2, sample))
    assert np.array_equal(H.compress_batch(sample), out[idx])     # position independence
    perm = rng.permutation(n)
    out2 = H.compress_batch(x.reshape(n, 2, f.n64)[perm])
    assert np.array_equal(out2, out[perm])                         # determinism under re-ordering
    # Merkle decomposition at scale: root(all) == root(roots of the 2^6 sub-trees)
    leaves = out[: 1 << 18]
    root = H.merkle_root(leaves)
    sub = np.concatenate([H.merkle_root(c) for c in leaves.reshape(64, -1, f.n64)])
    assert np.array_equal(H.merkle_root(sub), root)


@pytest.mark.parametrize("field,inst", CASES)
def test_adversarial_limb_patterns(field, inst):
    """States whose Montgomery limbs are all-ones / zero / sign-bit patterns (reduced mod p): long carry and
    borrow ripples that uniform inputs never produce. Compress and permute vs the oracle."""
    import random

    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f, W = H.FIELD, H.STATE_WIDTH
    rng = random.Random(99)
    special = [0, 1, 0x7FFFFFFFFFFFFFFF, 0x8000000000000000, 0xFFFFFFFFFFFFFFFE, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF00000000,
               0x00000000FFFFFFFF]
    vals = []
    for _ in range(W * 257):
        limbs = [rng.choice(special) if rng.random() < 0.8 else rng.getrandbits(64) for _ in range(f.n64)]
        v = sum(l << (64 * i) for i, l in enumerate(limbs)) % f.p
        vals.append(v)
    vals[:W] = [f.p - 1] * W
    vals[W:2 * W] = [0] * W
    x = np.array([[(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(f.n64)] for v in vals], dtype=np.uint64)
    assert np.array_equal(H.permutation_batch(x), C.permute(fi, ii, x))
    assert np.array_equal(H.compress_k_batch(x, W), C.compress(fi, ii, W, x))


def test_full_size_cross_mode_properties_4_3():
    """Size-independent identities on a 2^20-state Pallas Anemoi-4-3 batch (BASELINE config 3's node function):
    compress_k(.,4) = compress(.)[0] + compress(.)[1]  (anemoi_4_3/hasher.rs:148-179), and
    hash_field of a 3-element message = permutation([e0, e1, e2, 0])[0]  (hasher.rs:93-129, sigma = 1)."""
    H = A.AnemoiPallas_4_3
    f = H.FIELD
    n = 1 << 20
    x = f.random_mont(4 * n, SEED + 70)
    k4 = H.compress_k_batch(x, 4)
    k2 = H.compress_k_batch(x, 2)
    rng = np.random.default_rng(11)
    idx = np.sort(rng.choice(n, size=4096, replace=False))
    a = f.decode(k2.reshape(n, 2, f.n64)[idx, 0])
    b = f.decode(k2.reshape(n, 2, f.n64)[idx, 1])
    c = f.decode(k4[idx])
    assert all((u + v) % f.p == w for u, v, w in zip(a, b, c))
    # sponge vs permutation on 2^16 messages of 3 elements
    m = 1 << 16
    msgs = x[: 3 * m].reshape(m, 3, f.n64)
    states = np.zeros((m, 4, f.n64), dtype=np.uint64)
    states[:, :3] = msgs
    perm = H.permutation_batch(states.reshape(-1, f.n64)).reshape(m, 4, f.n64)
    assert np.array_equal(H.hash_field_batch(msgs), perm[:, 0])
    # 2-1: hash_field of a 1-element message = permutation([e, 0])[0]; of 2 elements = two chained permutations
    H2 = A.AnemoiBls12_381_2_1
    f2 = H2.FIELD
    e = f2.random_mont(1 << 14, SEED + 71)
    st = np.zeros((1 << 14, 2, f2.n64), dtype=np.uint64)
    st[:, 0] = e
    assert np.array_equal(H2.hash_field_batch(e.reshape(-1, 1, f2.n64)), H2.permutation_batch(st.reshape(-1, f2.n64)).reshape(-1, 2, f2.n64)[:, 0])


@pytest.mark.parametrize("field,inst", CASES)
def test_layers_in_isolation(field, inst):
    """Anemoi::{ark_layer, mds_layer, sbox_layer, round} one at a time on random states vs the oracle, and the
    composition ark -> mds -> sbox == round; round >= NUM_ROUNDS is rejected like the reference's assert."""
    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f, W = H.FIELD, H.STATE_WIDTH
    x = f.random_mont(W * 301, SEED + 90)
    for r in (0, 3, H.NUM_HASH_ROUNDS - 1):
        a = H.layer_batch(x, "ark", r)
        assert np.array_equal(a, C.layer(fi, ii, 0, r, x))
        m = H.layer_batch(a, "mds")
        assert np.array_equal(m, C.layer(fi, ii, 1, 0, a))
        s = H.layer_batch(m, "sbox")
        assert np.array_equal(s, C.layer(fi, ii, 2, 0, m))
        assert np.array_equal(H.layer_batch(x, "round", r), s)
        assert np.array_equal(s, C.layer(fi, ii, 3, r, x))
    with pytest.raises(A.LengthError):
        H.layer_batch(x, "round", H.NUM_HASH_ROUNDS)
    # the whole permutation = NUM_ROUNDS rounds + one more mds_layer (traits.rs:370-378)
    y = x[: W * 5].copy()
    for r in range(H.NUM_HASH_ROUNDS):
        y = H.layer_batch(y, "round", r)
    y = H.layer_batch(y, "mds")
    assert np.array_equal(y, H.permutation_batch(x[: W * 5]))


@pytest.mark.parametrize("field", ["bls12_381", "bn_254", "pallas"])
def test_count_noncanonical(field):
    """The opt-in input check counts exactly the elements >= p (p itself, p + 1, all-ones), none of the canonical ones."""
    H = HASHERS[(field, "anemoi_2_1")]
    f = H.FIELD
    x = f.random_mont(1000, SEED + 95)
    assert H.count_noncanonical(x) == 0
    bad = x.copy()
    mask = (1 << 64) - 1
    for row, v in ((3, f.p), (500, f.p + 1), (999, (1 << (64 * f.n64)) - 1)):
        for j in range(f.n64):
            bad[row, j] = (v >> (64 * j)) & mask
    bad[7] = [(((f.p - 1) >> (64 * j)) & mask) for j in range(f.n64)]   # p - 1 is canonical
    assert H.count_noncanonical(bad) == 3
    assert H.count_noncanonical(x[:0]) == 0


@pytest.mark.parametrize("field,inst", [("pallas", "anemoi_4_3"), ("ed_on_bls12_377", "anemoi_2_1"), ("bn_254", "anemoi_4_3"),
                                        ("bls12_377", "anemoi_4_3"), ("bls12_381", "anemoi_2_1")])
def test_latency_and_throughput_forms_agree(field, inst):
    """The same states through every launch geometry: the latency form (one warp per block, unchained carries; <= 2 warps
    per SM sub-partition), the throughput kernel with 32-thread blocks, and with 128-thread blocks -- identical limbs, and
    equal to the oracle on the common prefix. (For kernels whose throughput form is unchained the two forms are one kernel.)"""
    H = HASHERS[(field, inst)]
    fi, ii = ids(field, inst)
    f, W = H.FIELD, H.STATE_WIDTH
    cols = W // 2
    sizes = [100, 17000 // cols, 19500 // cols, 45000 // cols]   # lat, lat (many blocks), 32-thread throughput, 128-thread
    x = f.random_mont(sizes[-1] * W, SEED + 97)
    ref = C.compress(fi, ii, W, x[: 100 * W])
    outs = [H.compress_k_batch(x[: n * W], W) for n in sizes]
    for n, o in zip(sizes, outs):
        assert np.array_equal(o[:100], ref), "n = %d" % n
    for n, o in zip(sizes[:-1], outs[:-1]):
        assert np.array_equal(o, outs[-1][:n]), "n = %d differs from the largest batch" % n
    ffi.check(ffi.lib.anemoi_b200_pool_trim(0, 0))   # hand the cached device buffers back; the next call re-allocates
    assert np.array_equal(H.compress_k_batch(x[: 100 * W], W), ref)
    ffi.check(ffi.lib.anemoi_b200_pool_trim(0, 0))
    ffi.check(ffi.lib.anemoi_b200_pool_reserve(0, 64 << 20))   # grow the pool ahead of the call instead
    assert np.array_equal(H.compress_k_batch(x[: 100 * W], W), ref)
    assert ffi.lib.anemoi_b200_pool_reserve(0, 0) == 0
    assert ffi.lib.anemoi_b200_pool_reserve(1 << 20, 16) == ffi.ERR_ARG   # no such device
