"""Host-side mirror of the reference's public interface for the batched path.

One class per reference marker type (`AnemoiBls12_381_2_1`, `AnemoiPallas_4_3`, ... --
src/<field>/anemoi_{2_1,4_3}/mod.rs:38), each offering the reference's associated functions
  Sponge::{hash, hash_field, merge}      src/traits.rs:8-20
  Jive::{compress, compress_k}           src/traits.rs:23-33
  Anemoi::{permutation, sbox_layer}      src/traits.rs:328-378
with the same argument meaning and the same failure conditions (the reference's `assert!` panics
surface as AssertionError subclasses), plus `*_batch` forms that take whole batches. Every call goes
through the C ABI of libanemoi_b200.so; nothing is computed in Python.

Per-item calls take and return canonical Python ints (what `MontFp!("...")` literals denote).
Batch calls take and return numpy uint64 arrays of Montgomery limbs, shape (..., N64) -- the memory of
a Rust `&[Felt]` -- or CUDA torch tensors of dtype int64/uint64 with the same layout (then the `_dev`
entry points run on torch's current stream and nothing is copied).
"""
import ctypes

import numpy as np

from . import ffi
from .fields import FIELDS, INST_2_1, INST_4_3

_lib = ffi.lib


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _np_in(a, n64):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.size % n64:
        raise ffi.LengthError(ffi.ERR_LENGTH, "array is not a whole number of field elements")
    return a


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _torch_args(t):
    import torch

    if not t.is_cuda:
        raise TypeError("torch tensors passed to anemoi_rust_b200 must live on a CUDA device")
    if t.dtype not in (torch.int64, torch.uint64):
        raise TypeError("expected an int64/uint64 tensor of Montgomery limbs")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t


def _stream_of(t):
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _on_device_of(t):
    """The `_dev` entry points launch on the CURRENT device: make the tensor's device current around the call."""
    import torch

    return torch.cuda.device(t.device)


def _check_torch_out(out, like, numel):
    if out.device != like.device or out.dtype != like.dtype or not out.is_contiguous() or out.numel() != numel:
        raise ValueError("`out` must be a contiguous %s tensor of %d limbs on %s" % (like.dtype, numel, like.device))


def _check_np_out(out, numel):
    if not (isinstance(out, np.ndarray) and out.dtype == np.uint64 and out.flags.c_contiguous and out.flags.writeable
            and out.size == numel):
        raise ValueError("`out` must be a writable C-contiguous uint64 array of %d limbs" % numel)


class AnemoiDigest:
    """`AnemoiDigest([Felt; 1])` -- src/<field>/anemoi_*/digest.rs:13-53."""

    __slots__ = ("_e", "_field")

    def __init__(self, value, field):
        value = list(value)
        assert len(value) == 1  # DIGEST_SIZE
        self._e = [int(value[0]) % field.p]
        self._field = field

    @classmethod
    def new(cls, value, field):
        return cls(value, field)

    def as_elements(self):
        return self._e

    def to_elements(self):
        return list(self._e)

    @staticmethod
    def digests_to_elements(digests):
        out = []
        for d in digests:
            out.extend(d.as_elements())
        return out

    def to_bytes(self, device=0):
        """canonical little-endian bytes (digest.rs:42-46), de-Montgomery'd on the device."""
        f = self._field
        limbs = f.encode(self._e)
        out = np.empty(f.felt_bytes, dtype=np.uint8)
        ffi.check(_lib.anemoi_b200_digest_to_bytes(f.id, _ptr(limbs), _ptr(out), 1, device))
        return out.tobytes()

    def __eq__(self, other):
        return isinstance(other, AnemoiDigest) and self._e == other._e and self._field is other._field

    def __repr__(self):
        return "AnemoiDigest(%d)" % self._e[0]


class _AnemoiBase:
    FIELD = None      # fields.Field
    INST = None       # INST_2_1 / INST_4_3
    STATE_WIDTH = 0   # mod.rs:20
    RATE_WIDTH = 0    # mod.rs:22
    NUM_COLUMNS = 0   # mod.rs:25
    DIGEST_SIZE = 1   # mod.rs:28
    NUM_HASH_ROUNDS = 0
    device = 0        # device index used by host-pointer calls

    # ---- Digest helpers ------------------------------------------------------------------------
    @classmethod
    def Digest(cls, value):
        return AnemoiDigest(value, cls.FIELD)

    @classmethod
    def default_digest(cls):
        return AnemoiDigest([0], cls.FIELD)

    # ---- batched entry points (numpy host arrays or CUDA torch tensors) ------------------------
    @classmethod
    def permutation_batch(cls, states):
        """Anemoi::permutation on n states; returns a new array (numpy) or permutes in place (torch)."""
        f, W = cls.FIELD, cls.STATE_WIDTH
        if _is_torch(states):
            t = _torch_args(states)
            n = t.numel() // (W * f.n64)
            if t.numel() != n * W * f.n64:
                raise ffi.LengthError(ffi.ERR_LENGTH, "not a whole number of states")
            with _on_device_of(t):
                ffi.check(_lib.anemoi_b200_permute_dev(f.id, cls.INST, ctypes.c_void_p(t.data_ptr()), n, _stream_of(t)))
            return t
        a = _np_in(states, f.n64).copy()
        if a.size % (W * f.n64):
            raise ffi.LengthError(ffi.ERR_LENGTH, "not a whole number of states")
        ffi.check(_lib.anemoi_b200_permute(f.id, cls.INST, _ptr(a), a.size // (W * f.n64), cls.device))
        return a

    @classmethod
    def sbox_layer_batch(cls, states):
        f, W = cls.FIELD, cls.STATE_WIDTH
        a = _np_in(states, f.n64).copy()
        if a.size % (W * f.n64):
            raise ffi.LengthError(ffi.ERR_LENGTH, "not a whole number of states")
        ffi.check(_lib.anemoi_b200_sbox_layer(f.id, cls.INST, _ptr(a), a.size // (W * f.n64), cls.device))
        return a

    @classmethod
    def layer_batch(cls, states, layer, round_ctr=0):
        """One layer on n states (host array in, new array out): 'ark' (needs round_ctr), 'mds', 'sbox',
        'round' (needs round_ctr) -- Anemoi::{ark_layer, mds_layer, sbox_layer, round}, src/traits.rs:113-367."""
        f, W = cls.FIELD, cls.STATE_WIDTH
        code = {"ark": 0, "mds": 1, "sbox": 2, "round": 3}[layer]
        a = _np_in(states, f.n64).copy()
        if a.size % (W * f.n64):
            raise ffi.LengthError(ffi.ERR_LENGTH, "not a whole number of states")
        ffi.check(_lib.anemoi_b200_layer(f.id, cls.INST, code, round_ctr, _ptr(a), a.size // (W * f.n64), cls.device))
        return a

    @classmethod
    def compress_k_batch(cls, states, k, out=None, n_gpus=1):
        """Jive::compress_k on n states: (n*W felts) -> (n*W/k felts). n_gpus > 1 (host arrays only) splits the
        batch over that many devices of this process; independent states need no collective."""
        f, W = cls.FIELD, cls.STATE_WIDTH
        if k <= 0 or W % k or k % 2:
            raise ffi.ArityError(ffi.ERR_ARITY, "compress_k: k must be even and divide STATE_WIDTH")
        per = W // k
        if _is_torch(states):
            import torch

            t = _torch_args(states)
            n = t.numel() // (W * f.n64)
            if t.numel() != n * W * f.n64:
                raise ffi.LengthError(ffi.ERR_LENGTH, "not a whole number of states")
            if out is None:
                out = torch.empty((n * per, f.n64), dtype=t.dtype, device=t.device)
            else:
                _check_torch_out(out, t, n * per * f.n64)
            with _on_device_of(t):
                ffi.check(_lib.anemoi_b200_compress_dev(f.id, cls.INST, k, ctypes.c_void_p(t.data_ptr()),
                                                        ctypes.c_void_p(out.data_ptr()), n, _stream_of(t)))
            return out
        a = _np_in(states, f.n64)
        if a.size % (W * f.n64):
            raise ffi.LengthError(ffi.ERR_LENGTH, "not a whole number of states")
        n = a.size // (W * f.n64)
        if out is None:
            out = np.empty((n * per, f.n64), dtype=np.uint64)
        else:
            _check_np_out(out, n * per * f.n64)
        if n_gpus > 1:
            ffi.check(_lib.anemoi_b200_compress_multi(f.id, cls.INST, k, _ptr(a), _ptr(out), n, n_gpus))
        else:
            ffi.check(_lib.anemoi_b200_compress(f.id, cls.INST, k, _ptr(a), _ptr(out), n, cls.device))
        return out

    @classmethod
    def compress_batch(cls, states, out=None):
        return cls.compress_k_batch(states, 2, out)

    @classmethod
    def hash_field_batch(cls, elems, felts_per_msg=None, offsets=None, n_msgs=None):
        """Sponge::hash_field on many messages. Either elems has shape (n_msgs, L, N64) / felts_per_msg is
        given (fixed length; pass n_msgs too when felts_per_msg == 0), or `offsets` (n_msgs + 1 element
        offsets) describes ragged messages."""
        f = cls.FIELD
        if _is_torch(elems):
            import torch

            t = _torch_args(elems)
            if t.numel() % f.n64:
                raise ffi.LengthError(ffi.ERR_LENGTH, "tensor is not a whole number of field elements")
            if offsets is not None:
                o = _torch_args(offsets)
                if o.device != t.device:
                    raise ValueError("offsets and elems must live on the same device")
                n = o.numel() - 1
                out = torch.empty((n, f.n64), dtype=t.dtype, device=t.device)
                with _on_device_of(t):
                    ffi.check(_lib.anemoi_b200_hash_field_ragged_dev(
                        f.id, cls.INST, ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(o.data_ptr()), n,
                        ctypes.c_void_p(out.data_ptr()), _stream_of(t)))
                return out
            if felts_per_msg is None:
                assert t.dim() == 3
                felts_per_msg = t.shape[1]
            if felts_per_msg and t.numel() % (felts_per_msg * f.n64):
                raise ffi.LengthError(ffi.ERR_LENGTH, "not a whole number of messages of felts_per_msg elements")
            n = t.numel() // (felts_per_msg * f.n64) if felts_per_msg else (n_msgs or 0)
            out = torch.empty((n, f.n64), dtype=t.dtype, device=t.device)
            with _on_device_of(t):
                ffi.check(_lib.anemoi_b200_hash_field_dev(f.id, cls.INST, ctypes.c_void_p(t.data_ptr()), n, felts_per_msg,
                                                          ctypes.c_void_p(out.data_ptr()), _stream_of(t)))
            return out
        a = _np_in(elems, f.n64)
        if offsets is not None:
            o = np.ascontiguousarray(offsets, dtype=np.uint64)
            n = o.size - 1
            out = np.empty((n, f.n64), dtype=np.uint64)
            ffi.check(_lib.anemoi_b200_hash_field_ragged(f.id, cls.INST, _ptr(a), _ptr(o), n, _ptr(out), cls.device))
            return out
        if felts_per_msg is None:
            assert a.ndim == 3
            felts_per_msg = a.shape[1]
            n = a.shape[0]
        else:
            if felts_per_msg and a.size % (felts_per_msg * f.n64):
                raise ffi.LengthError(ffi.ERR_LENGTH, "not a whole number of messages of felts_per_msg elements")
            n = a.size // (felts_per_msg * f.n64) if felts_per_msg else (n_msgs or 0)
        out = np.empty((n, f.n64), dtype=np.uint64)
        if a.size == 0:
            a = np.zeros(f.n64, dtype=np.uint64)
        ffi.check(_lib.anemoi_b200_hash_field(f.id, cls.INST, _ptr(a), n, felts_per_msg, _ptr(out), cls.device))
        return out

    @classmethod
    def hash_batch(cls, data, bytes_per_msg=None):
        """Sponge::hash on many equal-length byte strings: data = bytes / uint8 array (n_msgs, L)."""
        f = cls.FIELD
        a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data,
                                 dtype=np.uint8)
        if bytes_per_msg is None:
            assert a.ndim == 2
            n, bytes_per_msg = a.shape
        else:
            if bytes_per_msg and a.size % bytes_per_msg:
                raise ffi.LengthError(ffi.ERR_LENGTH, "not a whole number of messages of bytes_per_msg bytes")
            n = a.size // bytes_per_msg if bytes_per_msg else 1
        out = np.empty((n, f.n64), dtype=np.uint64)
        ffi.check(_lib.anemoi_b200_hash_bytes(f.id, cls.INST, _ptr(a), n, bytes_per_msg, _ptr(out), cls.device))
        return out

    @classmethod
    def hash_ragged(cls, messages):
        """Sponge::hash on a list of byte strings of different lengths (one kernel launch)."""
        f = cls.FIELD
        lens = np.array([len(m) for m in messages], dtype=np.uint64)
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        blob = np.frombuffer(b"".join(bytes(m) for m in messages), dtype=np.uint8)
        if blob.size == 0:
            blob = np.zeros(1, dtype=np.uint8)
        out = np.empty((len(messages), f.n64), dtype=np.uint64)
        ffi.check(_lib.anemoi_b200_hash_bytes_ragged(f.id, cls.INST, _ptr(np.ascontiguousarray(blob)), _ptr(offs),
                                                     len(messages), _ptr(out), cls.device))
        return out

    @classmethod
    def merge_batch(cls, digest_pairs):
        """Sponge::merge on n pairs: (n, 2, N64) -> (n, N64)."""
        f = cls.FIELD
        a = _np_in(digest_pairs, f.n64)
        if a.size % (2 * f.n64):
            raise ffi.LengthError(ffi.ERR_LENGTH, "merge needs pairs of digests")
        n = a.size // (2 * f.n64)
        out = np.empty((n, f.n64), dtype=np.uint64)
        ffi.check(_lib.anemoi_b200_merge(f.id, cls.INST, _ptr(a), _ptr(out), n, cls.device))
        return out

    @classmethod
    def count_noncanonical(cls, elems):
        """Opt-in input check: how many elements of a host limb array are >= p (the compute entries assume none)."""
        f = cls.FIELD
        a = _np_in(elems, f.n64)
        out = np.zeros(1, dtype=np.uint64)
        ffi.check(_lib.anemoi_b200_count_noncanonical(f.id, _ptr(a), a.size // f.n64, _ptr(out), cls.device))
        return int(out[0])

    @classmethod
    def merkle_root(cls, leaves, n_gpus=1):
        """Jive Merkle root of arity STATE_WIDTH over a host array of leaf digests (n, N64)."""
        f = cls.FIELD
        a = _np_in(leaves, f.n64)
        out = np.empty((1, f.n64), dtype=np.uint64)
        if n_gpus > 1:
            # the library binds NCCL at run time and reuses a copy the process already mapped: make sure that copy is
            # PyTorch's (when there is one) rather than the system's, or a later `import torch` would trip over it
            try:
                import torch  # noqa: F401
            except ImportError:
                pass
        ffi.check(_lib.anemoi_b200_merkle_root(f.id, cls.INST, cls.STATE_WIDTH, _ptr(a), a.size // f.n64, _ptr(out),
                                               n_gpus))
        return out

    @classmethod
    def merkle_roots_batch(cls, leaves, leaves_per_tree):
        """Roots of many equal-size trees stored back to back (CUDA tensor in, CUDA tensor out): one launch
        per level over all trees at once."""
        from . import merkle

        height, m = 0, leaves_per_tree
        while m > 1:
            if m % cls.STATE_WIDTH:
                raise ffi.LengthError(ffi.ERR_LENGTH, "leaves_per_tree must be a power of the arity")
            m //= cls.STATE_WIDTH
            height += 1
        return merkle.merkle_reduce(cls, leaves, height)

    @classmethod
    def merkle_open(cls, leaves, indices):
        """Build the tree over host `leaves` (n, N64) and open `indices`: returns (root (1, N64),
        paths (n_idx, height, arity - 1, N64)). Path layout: siblings per level, leaf level first."""
        f, ar = cls.FIELD, cls.STATE_WIDTH
        a = _np_in(leaves, f.n64)
        n = a.size // f.n64
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        height, m = 0, n
        while m > 1:
            if m % ar:
                raise ffi.LengthError(ffi.ERR_LENGTH, "n_leaves must be a power of the arity")
            m //= ar
            height += 1
        root = np.empty((1, f.n64), dtype=np.uint64)
        paths = np.empty((idx.size, height, ar - 1, f.n64), dtype=np.uint64)
        ffi.check(_lib.anemoi_b200_merkle_open(f.id, cls.INST, ar, _ptr(a), n, _ptr(idx), idx.size, _ptr(root),
                                               _ptr(paths) if paths.size else None, cls.device))
        return root, paths

    @classmethod
    def merkle_verify(cls, leaf_values, indices, paths):
        """Recompute the root implied by each (leaf value, index, path); returns (n_idx, N64). The caller
        compares with the committed root."""
        f, ar = cls.FIELD, cls.STATE_WIDTH
        v = _np_in(leaf_values, f.n64)
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        pth = np.ascontiguousarray(paths, dtype=np.uint64)
        n_idx = idx.size
        height = pth.size // (n_idx * (ar - 1) * f.n64) if n_idx and pth.size else 0
        roots = np.empty((n_idx, f.n64), dtype=np.uint64)
        ffi.check(_lib.anemoi_b200_merkle_verify(f.id, cls.INST, ar, _ptr(v), _ptr(idx), _ptr(pth) if pth.size else None,
                                                 height, n_idx, _ptr(roots), cls.device))
        return roots

    # ---- the reference's per-item API (canonical ints) ------------------------------------------
    @classmethod
    def permutation(cls, state):
        """Anemoi::permutation(&mut state) -- src/traits.rs:370; mutates the list in place."""
        assert len(state) == cls.STATE_WIDTH  # debug_assert! traits.rs:371
        res = cls.FIELD.decode(cls.permutation_batch(cls.FIELD.encode(state)))
        state[:] = res

    @classmethod
    def sbox_layer(cls, state):
        assert len(state) == cls.STATE_WIDTH
        res = cls.FIELD.decode(cls.sbox_layer_batch(cls.FIELD.encode(state)))
        state[:] = res

    @classmethod
    def compress(cls, elems):
        """Jive::compress -- hasher.rs:96-103 / 4-3 :148-160."""
        if len(elems) != cls.STATE_WIDTH:  # assert!(elems.len() == STATE_WIDTH)
            raise ffi.LengthError(ffi.ERR_LENGTH, "compress: elems.len() != STATE_WIDTH")
        return cls.FIELD.decode(cls.compress_k_batch(cls.FIELD.encode(elems), 2))

    @classmethod
    def compress_k(cls, elems, k):
        """Jive::compress_k -- hasher.rs:105-110 / 4-3 :162-179."""
        if cls.INST == INST_2_1:
            if k != 2:  # assert!(k == 2)
                raise ffi.ArityError(ffi.ERR_ARITY, "compress_k: this instantiation only supports k = 2")
        if len(elems) != cls.STATE_WIDTH:
            raise ffi.LengthError(ffi.ERR_LENGTH, "compress_k: elems.len() != STATE_WIDTH")
        return cls.FIELD.decode(cls.compress_k_batch(cls.FIELD.encode(elems), k))

    @classmethod
    def hash_field(cls, elems):
        """Sponge::hash_field -- hasher.rs:68-85 / 4-3 :93-129."""
        f = cls.FIELD
        elems = list(elems)
        enc = f.encode(elems) if elems else np.zeros((1, f.n64), dtype=np.uint64)
        out = np.empty((1, f.n64), dtype=np.uint64)
        ffi.check(_lib.anemoi_b200_hash_field(f.id, cls.INST, _ptr(enc), 1, len(elems), _ptr(out), cls.device))
        return AnemoiDigest(f.decode(out), f)

    @classmethod
    def hash(cls, data):
        """Sponge::hash(bytes) -- hasher.rs:18-66 / 4-3 :18-91."""
        f = cls.FIELD
        a = np.frombuffer(bytes(data), dtype=np.uint8)
        buf = a if a.size else np.zeros(1, dtype=np.uint8)
        out = np.empty((1, f.n64), dtype=np.uint64)
        ffi.check(_lib.anemoi_b200_hash_bytes(f.id, cls.INST, _ptr(np.ascontiguousarray(buf)), 1, a.size, _ptr(out),
                                              cls.device))
        return AnemoiDigest(f.decode(out), f)

    @classmethod
    def merge(cls, digests):
        """Sponge::merge(&[Digest; 2]) -- 2-1: Jive (hasher.rs:87-92); 4-3: sponge, reads digests[0] only
        (anemoi_4_3/hasher.rs:131-144, reproduced as written)."""
        assert len(digests) == 2
        elems = AnemoiDigest.digests_to_elements(digests)
        out = cls.merge_batch(cls.FIELD.encode(elems))
        return AnemoiDigest(cls.FIELD.decode(out), cls.FIELD)


def _make(field_name, inst, rust_name):
    f = FIELDS[field_name]
    rounds = ffi.lib.anemoi_b200_num_rounds(f.id, inst)
    attrs = dict(FIELD=f, INST=inst, STATE_WIDTH=2 if inst == INST_2_1 else 4, RATE_WIDTH=1 if inst == INST_2_1 else 3,
                 NUM_COLUMNS=1 if inst == INST_2_1 else 2, NUM_HASH_ROUNDS=rounds,
                 __doc__="%s -- src/%s/anemoi_%s/mod.rs:38" % (rust_name, field_name, "2_1" if inst == INST_2_1 else "4_3"))
    return type(rust_name, (_AnemoiBase,), attrs)


# struct names exactly as in the reference
_RUST_NAMES = {
    "bls12_377": "AnemoiBls12_377", "bls12_381": "AnemoiBls12_381", "bn_254": "AnemoiBn254",
    "ed_on_bls12_377": "AnemoiEdOnBls12_377", "jubjub": "AnemoiJubjub", "pallas": "AnemoiPallas", "vesta": "AnemoiVesta",
}
HASHERS = {}
for _fname, _rname in _RUST_NAMES.items():
    for _inst, _suffix in ((INST_2_1, "_2_1"), (INST_4_3, "_4_3")):
        _cls = _make(_fname, _inst, _rname + _suffix)
        globals()[_rname + _suffix] = _cls
        HASHERS[(_fname, "anemoi" + _suffix)] = _cls

__all__ = ["AnemoiDigest", "HASHERS"] + [c.__name__ for c in HASHERS.values()]
