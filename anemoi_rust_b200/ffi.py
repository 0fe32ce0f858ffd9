"""ctypes binding of libanemoi_b200.so (include/anemoi_b200.h). Plumbing only: every hash is computed by
the CUDA kernels behind the C ABI. Importing this module fails loudly when the library has not been
built (python __graft_entry__.py build / make); there is no Python or CPU fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ANEMOI_B200_LIB: developer override used to A/B-test alternative BUILDS of the same library (tools/ab_variants.sh)
LIB_PATH = os.environ.get("ANEMOI_B200_LIB") or os.path.join(_HERE, "libanemoi_b200.so")

OK = 0
ERR_ARG, ERR_FIELD, ERR_INST, ERR_ARITY, ERR_LENGTH, ERR_CUDA, ERR_NO_DEVICE, ERR_NOMEM, ERR_NCCL = -1, -2, -3, -4, -5, -6, -7, -8, -9


class AnemoiError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("anemoi_b200 error %d: %s" % (code, text))
        self.code = code


class ArityError(AnemoiError, AssertionError):
    """The reference `assert!`s on these (hasher.rs:97,107; 4-3 :149,163-165): a panic there, this here."""


class LengthError(AnemoiError, AssertionError):
    pass


class NoDeviceError(AnemoiError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "%s is missing: build it with `python __graft_entry__.py build` (or `make`). "
        "anemoi_rust_b200 has no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
_SIGS = {
    "anemoi_b200_version": ([], _i),
    "anemoi_b200_strerror": ([_i], ctypes.c_char_p),
    "anemoi_b200_last_cuda_error": ([], ctypes.c_char_p),
    "anemoi_b200_device_count": ([], _i),
    "anemoi_b200_field_limbs": ([_i], _i),
    "anemoi_b200_state_width": ([_i], _i),
    "anemoi_b200_rate_width": ([_i], _i),
    "anemoi_b200_num_rounds": ([_i, _i], _i),
    "anemoi_b200_field_name": ([_i], ctypes.c_char_p),
    "anemoi_b200_permute": ([_i, _i, _vp, _sz, _i], _i),
    "anemoi_b200_sbox_layer": ([_i, _i, _vp, _sz, _i], _i),
    "anemoi_b200_layer": ([_i, _i, _i, _i, _vp, _sz, _i], _i),
    "anemoi_b200_layer_dev": ([_i, _i, _i, _i, _vp, _sz, _vp], _i),
    "anemoi_b200_compress": ([_i, _i, _i, _vp, _vp, _sz, _i], _i),
    "anemoi_b200_compress_multi": ([_i, _i, _i, _vp, _vp, _sz, _i], _i),
    "anemoi_b200_hash_field": ([_i, _i, _vp, _sz, _sz, _vp, _i], _i),
    "anemoi_b200_hash_field_ragged": ([_i, _i, _vp, _vp, _sz, _vp, _i], _i),
    "anemoi_b200_hash_bytes": ([_i, _i, _vp, _sz, _sz, _vp, _i], _i),
    "anemoi_b200_hash_bytes_ragged": ([_i, _i, _vp, _vp, _sz, _vp, _i], _i),
    "anemoi_b200_hash_bytes_ragged_dev": ([_i, _i, _vp, _vp, _sz, _vp, _vp], _i),
    "anemoi_b200_merge": ([_i, _i, _vp, _vp, _sz, _i], _i),
    "anemoi_b200_merkle_root": ([_i, _i, _i, _vp, _sz, _vp, _i], _i),
    "anemoi_b200_digest_to_bytes": ([_i, _vp, _vp, _sz, _i], _i),
    "anemoi_b200_permute_dev": ([_i, _i, _vp, _sz, _vp], _i),
    "anemoi_b200_sbox_layer_dev": ([_i, _i, _vp, _sz, _vp], _i),
    "anemoi_b200_compress_dev": ([_i, _i, _i, _vp, _vp, _sz, _vp], _i),
    "anemoi_b200_hash_field_dev": ([_i, _i, _vp, _sz, _sz, _vp, _vp], _i),
    "anemoi_b200_hash_field_ragged_dev": ([_i, _i, _vp, _vp, _sz, _vp, _vp], _i),
    "anemoi_b200_hash_bytes_dev": ([_i, _i, _vp, _sz, _sz, _vp, _vp], _i),
    "anemoi_b200_merge_dev": ([_i, _i, _vp, _vp, _sz, _vp], _i),
    "anemoi_b200_digest_to_bytes_dev": ([_i, _vp, _vp, _sz, _vp], _i),
    "anemoi_b200_merkle_reduce_dev": ([_i, _i, _i, _vp, _sz, _i, _vp, _vp, _vp], _i),
    "anemoi_b200_merkle_scratch_felts": ([_i, _sz], _sz),
    "anemoi_b200_merkle_root_sharded_dev": ([_i, _i, _i, _vp, _sz, _vp, _vp, _vp, _vp], _i),
    "anemoi_b200_merkle_sharded_scratch_felts": ([_i, _sz, _i], _sz),
    "anemoi_b200_nccl_version": ([], _i),
    "anemoi_b200_comm_unique_id": ([_vp], _i),
    "anemoi_b200_comm_init_rank": ([_vp, _i, _i, ctypes.POINTER(ctypes.c_void_p)], _i),
    "anemoi_b200_comm_info": ([_vp, ctypes.POINTER(_i), ctypes.POINTER(_i)], _i),
    "anemoi_b200_comm_destroy": ([_vp], _i),
    "anemoi_b200_pool_trim": ([_i, _sz], _i),
    "anemoi_b200_pool_reserve": ([_i, _sz], _i),
    "anemoi_b200_count_noncanonical": ([_i, _vp, _sz, _vp, _i], _i),
    "anemoi_b200_count_noncanonical_dev": ([_i, _vp, _sz, _vp, _vp], _i),
    "anemoi_b200_merkle_tree_felts": ([_i, _sz], _sz),
    "anemoi_b200_merkle_tree_dev": ([_i, _i, _i, _vp, _sz, _vp, _vp], _i),
    "anemoi_b200_merkle_open_dev": ([_i, _i, _i, _vp, _vp, _sz, _vp, _sz, _vp, _vp], _i),
    "anemoi_b200_merkle_verify_dev": ([_i, _i, _i, _vp, _vp, _vp, _i, _sz, _vp, _vp, _vp], _i),
    "anemoi_b200_merkle_open": ([_i, _i, _i, _vp, _sz, _vp, _sz, _vp, _vp, _i], _i),
    "anemoi_b200_merkle_verify": ([_i, _i, _i, _vp, _vp, _vp, _i, _sz, _vp, _i], _i),
    "anemoi_b200_imad_peak": ([_i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)], _i),
}
EXPORTED = sorted(_SIGS)
for _name, (_args, _res) in _SIGS.items():
    _f = getattr(lib, _name)  # AttributeError here = the .so does not export what the header declares
    _f.argtypes = _args
    _f.restype = _res


def check(rc):
    """Map a C-ABI status to the exception the reference's panic corresponds to."""
    if rc == OK:
        return
    text = lib.anemoi_b200_strerror(rc).decode()
    if rc in (ERR_CUDA, ERR_NO_DEVICE, ERR_NOMEM, ERR_NCCL):
        detail = lib.anemoi_b200_last_cuda_error().decode()
        if detail:
            text += " [" + detail + "]"
    if rc == ERR_ARITY:
        raise ArityError(rc, text)
    if rc == ERR_LENGTH:
        raise LengthError(rc, text)
    if rc == ERR_NO_DEVICE:
        raise NoDeviceError(rc, text)
    raise AnemoiError(rc, text)


def imad_peak(variant=3):
    ops, mhz = ctypes.c_double(), ctypes.c_double()
    check(lib.anemoi_b200_imad_peak(variant, ctypes.byref(ops), ctypes.byref(mhz)))
    return ops.value, mhz.value
