"""ctypes loader for the C oracle (oracle/anemoi_oracle.c -> oracle/libanemoi_oracle.so).
TEST INFRASTRUCTURE: importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libanemoi_oracle.so")
FIELDS = ["bls12_377", "bls12_381", "bn_254", "ed_on_bls12_377", "jubjub", "pallas", "vesta"]
N64 = [6, 6, 4, 4, 4, 4, 4]


def build(force=False):
    src = os.path.join(_HERE, "anemoi_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "params_gen.h"))):
        subprocess.check_call(["gcc", "-O3", "-funroll-loops", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        vp, i, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        _lib.oracle_permute.argtypes = [i, i, vp, sz]
        _lib.oracle_sbox_layer.argtypes = [i, i, vp, sz]
        _lib.oracle_compress.argtypes = [i, i, i, vp, vp, sz]
        _lib.oracle_layer.argtypes = [i, i, i, i, vp, sz]
        _lib.oracle_hash_field.argtypes = [i, i, vp, sz, sz, vp]
        _lib.oracle_hash_field_ragged.argtypes = [i, i, vp, vp, sz, vp]
        _lib.oracle_hash_bytes.argtypes = [i, i, vp, sz, sz, vp]
        _lib.oracle_merge.argtypes = [i, i, vp, vp, sz]
        _lib.oracle_merkle_root.argtypes = [i, i, i, vp, sz, vp, vp]
        _lib.oracle_digest_to_bytes.argtypes = [i, vp, vp, sz]
        _lib.oracle_set_threads.argtypes = [i]
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def _chk(rc):
    if rc != 0:
        raise AssertionError("oracle rc=%d" % rc)


def set_threads(t):
    lib().oracle_set_threads(t)


def max_threads():
    return lib().oracle_max_threads()


def permute(fi, inst, states):
    a = np.ascontiguousarray(states, dtype=np.uint64).copy()
    W = 2 if inst == 0 else 4
    _chk(lib().oracle_permute(fi, inst, _p(a), a.size // (W * N64[fi])))
    return a


def sbox_layer(fi, inst, states):
    a = np.ascontiguousarray(states, dtype=np.uint64).copy()
    W = 2 if inst == 0 else 4
    _chk(lib().oracle_sbox_layer(fi, inst, _p(a), a.size // (W * N64[fi])))
    return a


def layer(fi, inst, which, round_ctr, states):
    a = np.ascontiguousarray(states, dtype=np.uint64).copy()
    W = 2 if inst == 0 else 4
    _chk(lib().oracle_layer(fi, inst, which, round_ctr, _p(a), a.size // (W * N64[fi])))
    return a


def compress(fi, inst, k, states):
    a = np.ascontiguousarray(states, dtype=np.uint64)
    W = 2 if inst == 0 else 4
    n = a.size // (W * N64[fi])
    out = np.empty((n * (W // k), N64[fi]), dtype=np.uint64)
    _chk(lib().oracle_compress(fi, inst, k, _p(a), _p(out), n))
    return out


def hash_field(fi, inst, elems, n_msgs, length):
    a = np.ascontiguousarray(elems, dtype=np.uint64)
    if a.size == 0:
        a = np.zeros(N64[fi], dtype=np.uint64)
    out = np.empty((n_msgs, N64[fi]), dtype=np.uint64)
    _chk(lib().oracle_hash_field(fi, inst, _p(a), n_msgs, length, _p(out)))
    return out


def hash_field_ragged(fi, inst, elems, offsets):
    a = np.ascontiguousarray(elems, dtype=np.uint64)
    if a.size == 0:
        a = np.zeros(N64[fi], dtype=np.uint64)
    o = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = np.empty((o.size - 1, N64[fi]), dtype=np.uint64)
    _chk(lib().oracle_hash_field_ragged(fi, inst, _p(a), _p(o), o.size - 1, _p(out)))
    return out


def hash_bytes(fi, inst, data, n_msgs, nbytes):
    a = np.ascontiguousarray(data, dtype=np.uint8)
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint8)
    out = np.empty((n_msgs, N64[fi]), dtype=np.uint64)
    _chk(lib().oracle_hash_bytes(fi, inst, _p(a), n_msgs, nbytes, _p(out)))
    return out


def merge(fi, inst, pairs):
    a = np.ascontiguousarray(pairs, dtype=np.uint64)
    n = a.size // (2 * N64[fi])
    out = np.empty((n, N64[fi]), dtype=np.uint64)
    _chk(lib().oracle_merge(fi, inst, _p(a), _p(out), n))
    return out


def merkle_root(fi, inst, arity, leaves):
    a = np.ascontiguousarray(leaves, dtype=np.uint64)
    n = a.size // N64[fi]
    scratch = np.empty(((n // arity) + (n // arity) // arity + 2) * N64[fi], dtype=np.uint64)
    out = np.empty((1, N64[fi]), dtype=np.uint64)
    _chk(lib().oracle_merkle_root(fi, inst, arity, _p(a), n, _p(scratch), _p(out)))
    return out


def digest_to_bytes(fi, digests):
    a = np.ascontiguousarray(digests, dtype=np.uint64)
    n = a.size // N64[fi]
    out = np.empty(n * N64[fi] * 8, dtype=np.uint8)
    _chk(lib().oracle_digest_to_bytes(fi, _p(a), _p(out), n))
    return out
