"""Pin the C oracle (oracle/anemoi_oracle.c, the CPU restatement that the GPU is checked against and that
bench.py times as the CPU baseline): all 420 reference KATs, plus agreement with the big-integer oracle on
random inputs, limb for limb in Montgomery form."""
import random

import numpy as np
import pytest

from oracle import anemoi_ref as R
from oracle import c_oracle as C
from anemoi_rust_b200.fields import FIELDS as FLD

CASES = [(f, i) for f in R.FIELDS for i in R.INSTS]


def ints(v):
    if isinstance(v, list):
        return [ints(x) for x in v]
    return int(v)


def enc(field, vals):
    return FLD[field].encode([int(v) for v in vals])


def dec(field, limbs):
    return FLD[field].decode(limbs)


@pytest.mark.parametrize("field,inst", CASES)
def test_kats(kat, field, inst):
    fi, ii = R.FIELDS.index(field), R.INSTS.index(inst)
    k = kat[field][inst]
    for i, o in zip(ints(k["sbox"]["in"]), ints(k["sbox"]["out"])):
        assert dec(field, C.sbox_layer(fi, ii, enc(field, i))) == o
    for i, o in zip(ints(k["hash_field"]["in"]), ints(k["hash_field"]["out"])):
        assert dec(field, C.hash_field(fi, ii, enc(field, i), 1, len(i))) == [o]
    for h, o in zip(k["hash_bytes"]["in_hex"], ints(k["hash_bytes"]["out"])):
        b = np.frombuffer(bytes.fromhex(h), dtype=np.uint8)
        assert dec(field, C.hash_bytes(fi, ii, b, 1, b.size)) == [o]
    for i, o in zip(ints(k["jive2"]["in"]), ints(k["jive2"]["out"])):
        assert dec(field, C.compress(fi, ii, 2, enc(field, i))) == o
        if inst == "anemoi_2_1":
            assert dec(field, C.merge(fi, ii, enc(field, i))) == o
    if inst == "anemoi_4_3":
        for i, o in zip(ints(k["jive4"]["in"]), ints(k["jive4"]["out"])):
            assert dec(field, C.compress(fi, ii, 4, enc(field, i))) == o


@pytest.mark.parametrize("field,inst", CASES)
def test_random_vs_bigint(field, inst):
    fi, ii = R.FIELDS.index(field), R.INSTS.index(inst)
    P = R.params(field, inst)
    rng = random.Random(99 + 7 * fi + ii)
    W = P.width
    states = [[rng.randrange(P.p) for _ in range(W)] for _ in range(4)]
    flat = [v for s in states for v in s]
    got = dec(field, C.permute(fi, ii, enc(field, flat)))
    exp = []
    for s in states:
        t = list(s)
        R.permutation(P, t)
        exp += t
    assert got == exp
    for k in ([2] if W == 2 else [2, 4]):
        got = dec(field, C.compress(fi, ii, k, enc(field, flat)))
        exp = [v for s in states for v in R.compress_k(P, s, k)]
        assert got == exp
    for L in (0, 1, 2, 3, 4, 7):
        msg = [rng.randrange(P.p) for _ in range(L)]
        assert dec(field, C.hash_field(fi, ii, enc(field, msg), 1, L)) == [R.hash_field(P, msg)]
    for nb in (0, 1, 30, 31, 32, 46, 47, 48, 93, 94, 95, 200):
        data = bytes(rng.randrange(256) for _ in range(nb))
        assert dec(field, C.hash_bytes(fi, ii, np.frombuffer(data, dtype=np.uint8), 1, nb)) == [R.hash_bytes(P, data)]
    d = [rng.randrange(P.p) for _ in range(2)]
    assert dec(field, C.merge(fi, ii, enc(field, d))) == [R.merge(P, d[0], d[1])]
    leaves = [rng.randrange(P.p) for _ in range(W * W)]
    assert dec(field, C.merkle_root(fi, ii, W, enc(field, leaves))) == [R.merkle_root(P, leaves, W)]
    assert bytes(C.digest_to_bytes(fi, enc(field, d[:1]))) == R.digest_to_bytes(P, d[0])


def test_ragged_and_threads():
    fi, ii = 2, 1
    P = R.params("bn_254", "anemoi_4_3")
    rng = random.Random(5)
    lens = [0, 1, 2, 3, 4, 5, 6, 9]
    msgs = [[rng.randrange(P.p) for _ in range(L)] for L in lens]
    flat = [v for m in msgs for v in m]
    offs = np.cumsum([0] + lens).astype(np.uint64)
    got = dec("bn_254", C.hash_field_ragged(fi, ii, enc("bn_254", flat), offs))
    assert got == [R.hash_field(P, m) for m in msgs]
    assert C.max_threads() >= 1


@pytest.mark.parametrize("field,inst", CASES)
def test_layers_vs_bigint(field, inst):
    """ark_layer / mds_layer / round in isolation: C oracle == big-integer oracle (they are pinned by the
    reference's vectors only transitively, through hash/jive)."""
    fi, ii = R.FIELDS.index(field), R.INSTS.index(inst)
    P = R.params(field, inst)
    rng = random.Random(7 + fi)
    s0 = [rng.randrange(P.p) for _ in range(P.width)]
    for r in (0, 1, P.rounds - 1):
        s = list(s0)
        R.ark(P, s, r)
        assert dec(field, C.layer(fi, ii, 0, r, enc(field, s0))) == s
        s = list(s0)
        R.ark(P, s, r)
        R.mds(P, s)
        R.sbox(P, s)
        assert dec(field, C.layer(fi, ii, 3, r, enc(field, s0))) == s
    s = list(s0)
    R.mds(P, s)
    assert dec(field, C.layer(fi, ii, 1, 0, enc(field, s0))) == s
