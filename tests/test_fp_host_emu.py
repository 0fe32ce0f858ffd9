"""The device Montgomery templates (anemoi_rust_b200/csrc/fp.cuh), compiled for the host with the PTX
carry-flag primitives emulated, checked against Python big integers: mul, sqr (canonical and lazy
[0,2p) variants), add, sub, beta-multiple, on random and boundary operands for all 7 fields."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import anemoi_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("fp_emu") / "libfp_emu.so"
    src = os.path.join(ROOT, "tests", "host_emu", "fp_emu.cpp")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(out), src])
    lib = ctypes.CDLL(str(out))
    lib.fp_emu_op.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.fp_emu_op_flipped.argtypes = lib.fp_emu_op.argtypes
    return lib


def call(lib, fi, op, n, a, b=0):
    A = (ctypes.c_uint32 * n)(*[(a >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
    B = (ctypes.c_uint32 * n)(*[(b >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
    Rr = (ctypes.c_uint32 * n)()
    assert lib.fp_emu_op(fi, op, A, B, Rr) == 0
    res = sum(int(Rr[i]) << (32 * i) for i in range(n))
    # the same operation with the code-generation switch (carry-chained fix-ups) flipped must give the identical limbs:
    # both paths are live, chosen per kernel by measurement
    R2 = (ctypes.c_uint32 * n)()
    assert lib.fp_emu_op_flipped(fi, op, A, B, R2) == 0
    assert list(R2) == list(Rr), "chained / unchained carry paths disagree"
    return res


@pytest.mark.parametrize("field", R.FIELDS)
def test_fp_ops(emu, field):
    P = R.params(field, "anemoi_2_1")
    fi = R.FIELDS.index(field)
    n = emu.fp_emu_limbs(fi)
    assert n == 2 * P.n64
    p = P.p
    Rm = 1 << (32 * n)
    Rinv = pow(Rm, -1, p)
    spare = 32 * n - p.bit_length()
    rng = random.Random(1234 + fi)
    edge = [0, 1, 2, p - 1, p - 2, Rm % p, (p - 1) // 2, (p + 1) // 2, (1 << (p.bit_length() - 1)), (1 << 32) - 1,
            ((1 << (p.bit_length() - 1)) - 1)]
    vals = edge + [rng.randrange(p) for _ in range(60)]
    for a in vals:
        for b in rng.sample(vals, 6) + [a]:
            assert call(emu, fi, 0, n, a, b) == a * b * Rinv % p
            assert call(emu, fi, 4, n, a, b) == (a + b) % p
            assert call(emu, fi, 5, n, a, b) == (a - b) % p
        assert call(emu, fi, 1, n, a) == a * a * Rinv % p
        assert call(emu, fi, 6, n, a) == P.beta * a % p
    # lazy variants: >= 2 spare bits -> closed on [0, 2p); Pallas/Vesta (4p = R + tiny) -> drift < 2^127 per multiply
    near_lazy = spare < 2 and 0 <= 4 * p - Rm < (1 << 130)
    if spare >= 2 or near_lazy:
        drift = (1 << 140) if near_lazy else 0
        bound_in = 2 * p + drift
        lazy = vals + [2 * p - 1, 2 * p - 2, p, p + 1, bound_in - 1] + [rng.randrange(bound_in) for _ in range(60)]
        for a in lazy:
            for b in rng.sample(lazy, 4) + [a, bound_in - 1]:
                r = call(emu, fi, 2, n, a, b)
                assert r % p == a * b * Rinv % p
                assert r < 2 * p + (drift + (1 << 127) if near_lazy else 0) + (0 if near_lazy else 0) or r < 2 * p
            r = call(emu, fi, 3, n, a)
            assert r % p == a * a * Rinv % p
            assert r < (2 * p + 2 * drift + (1 << 128) if near_lazy else 2 * p)


def adversarial_values(p, n, rng, count=200):
    """Operands whose limbs are drawn from {0, 1, 0x7fffffff, 0x80000000, 0xfffffffe, 0xffffffff, random}:
    long carry/borrow ripples and all-ones products that uniform sampling never hits."""
    special = [0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF]
    out = []
    while len(out) < count:
        limbs = [rng.choice(special) if rng.random() < 0.8 else rng.getrandbits(32) for _ in range(n)]
        v = sum(l << (32 * i) for i, l in enumerate(limbs))
        out.append(v % p)
        out.append((v >> rng.randrange(0, 64)) % p)
    return out


@pytest.mark.parametrize("field", R.FIELDS)
def test_fp_adversarial_limbs(emu, field):
    P = R.params(field, "anemoi_2_1")
    fi = R.FIELDS.index(field)
    n = emu.fp_emu_limbs(fi)
    p = P.p
    Rinv = pow(1 << (32 * n), -1, p)
    rng = random.Random(4242 + fi)
    vals = adversarial_values(p, n, rng)
    for i, a in enumerate(vals):
        b = vals[(7 * i + 3) % len(vals)]
        assert call(emu, fi, 0, n, a, b) == a * b * Rinv % p
        assert call(emu, fi, 1, n, a) == a * a * Rinv % p
        assert call(emu, fi, 4, n, a, b) == (a + b) % p
        assert call(emu, fi, 5, n, a, b) == (a - b) % p
    spare = 32 * n - p.bit_length()
    near_lazy = spare < 2 and 0 <= 4 * p - (1 << (32 * n)) < (1 << 130)
    if spare >= 2 or near_lazy:
        hi = 2 * p + ((1 << 140) if near_lazy else 0)
        lazy = [v + rng.choice([0, p, hi - p - 1]) for v in vals]
        lazy = [v if v < hi else v - p for v in lazy]
        for i, a in enumerate(lazy):
            b = lazy[(5 * i + 1) % len(lazy)]
            assert call(emu, fi, 2, n, a, b) % p == a * b * Rinv % p
            assert call(emu, fi, 3, n, a) % p == a * a * Rinv % p
