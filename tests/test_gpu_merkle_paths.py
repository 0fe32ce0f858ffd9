"""GPU parity for SURVEY.md 8(f3): retained trees, authentication paths, verification, batches of trees.
The oracle side (tests only) builds every level with the C oracle's compress and derives the expected
siblings from it."""
import numpy as np
import pytest

import anemoi_rust_b200 as A
from oracle import c_oracle as C

pytestmark = pytest.mark.gpu


def oracle_levels(fi, ii, arity, leaves, n64):
    levels = [leaves.reshape(-1, n64)]
    while levels[-1].shape[0] > 1:
        levels.append(C.compress(fi, ii, arity, levels[-1]).reshape(-1, n64))
    return levels


def oracle_path(levels, arity, index):
    out = []
    idx = index
    for lvl in levels[:-1]:
        pos = idx % arity
        base = idx - pos
        out.append(np.stack([lvl[base + k] for k in range(arity) if k != pos]))
        idx //= arity
    return np.stack(out) if out else np.zeros((0, arity - 1, levels[0].shape[1]), dtype=np.uint64)


@pytest.mark.parametrize("field,inst,height", [("bls12_381", "anemoi_2_1", 7), ("bls12_377", "anemoi_2_1", 3),
                                               ("pallas", "anemoi_4_3", 4), ("vesta", "anemoi_4_3", 2),
                                               ("bn_254", "anemoi_4_3", 3), ("jubjub", "anemoi_2_1", 5)])
def test_open_and_verify(field, inst, height):
    H = A.HASHERS[(field, inst)]
    f, ar = H.FIELD, H.STATE_WIDTH
    fi, ii = A.FIELD_NAMES.index(field), (0 if inst == "anemoi_2_1" else 1)
    n = ar ** height
    leaves = f.random_mont(n, 0xA7E301 + height)
    levels = oracle_levels(fi, ii, ar, leaves, f.n64)
    rng = np.random.default_rng(3)
    idx = np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, size=17)])).astype(np.uint64)
    root, paths = H.merkle_open(leaves, idx)
    assert np.array_equal(root, levels[-1])
    assert paths.shape == (idx.size, height, ar - 1, f.n64)
    for q, i in enumerate(idx):
        assert np.array_equal(paths[q], oracle_path(levels, ar, int(i)))
    roots = H.merkle_verify(leaves[idx.astype(np.int64)], idx, paths)
    assert np.array_equal(roots, np.repeat(root, idx.size, axis=0))
    # a corrupted sibling or a wrong index must not verify
    bad = paths.copy()
    bad[0, height - 1, 0, 0] ^= np.uint64(1)
    assert not np.array_equal(H.merkle_verify(leaves[idx.astype(np.int64)], idx, bad)[0], root[0])
    if n > 1:
        wrong = idx.copy()
        wrong[0] = (wrong[0] + 1) % n
        assert not np.array_equal(H.merkle_verify(leaves[idx.astype(np.int64)], wrong, paths)[0], root[0])


def test_single_leaf_tree_and_errors():
    H = A.AnemoiPallas_4_3
    f = H.FIELD
    leaf = f.random_mont(1, 9)
    root, paths = H.merkle_open(leaf, [0])
    assert np.array_equal(root, leaf) and paths.size == 0
    with pytest.raises(A.LengthError):
        H.merkle_open(f.random_mont(12, 1), [0])
    with pytest.raises(A.AnemoiError):
        H.merkle_open(f.random_mont(16, 1), [16])   # index out of range


def test_batch_of_trees_and_device_tree():
    import ctypes

    import torch

    from anemoi_rust_b200 import ffi

    H = A.AnemoiVesta_4_3
    f = H.FIELD
    fi, ii = A.FIELD_NAMES.index("vesta"), 1
    n_trees, per = 37, 4 ** 3
    leaves = f.random_mont(n_trees * per, 0xBEEF)
    t = torch.from_numpy(leaves.view(np.int64)).cuda()
    roots = H.merkle_roots_batch(t, per).cpu().numpy().view(np.uint64)
    exp = np.concatenate([C.merkle_root(fi, ii, 4, c) for c in leaves.reshape(n_trees, per, f.n64)])
    assert np.array_equal(roots, exp)
    # retained tree on device: last element is the root, first level equals compress of the leaves
    one = t[:per].contiguous()
    tree = torch.empty((ffi.lib.anemoi_b200_merkle_tree_felts(4, per), f.n64), dtype=torch.int64, device="cuda")
    assert tree.shape[0] == 16 + 4 + 1
    ffi.check(ffi.lib.anemoi_b200_merkle_tree_dev(f.id, 1, 4, ctypes.c_void_p(one.data_ptr()), per,
                                                  ctypes.c_void_p(tree.data_ptr()), None))
    torch.cuda.synchronize()
    tr = tree.cpu().numpy().view(np.uint64)
    assert np.array_equal(tr[-1:], exp[:1])
    assert np.array_equal(tr[:16], C.compress(fi, ii, 4, leaves[:per]))
