"""anemoi_rust_b200 -- B200-native batched Anemoi engine (drop-in for the batched path of
anemoi-hash/anemoi-rust). The product is libanemoi_b200.so (hand-written sm_100a CUDA behind the C ABI
of include/anemoi_b200.h); this package is the host-side mirror of the reference's interface plus the
torch.distributed plumbing for multi-GPU Merkle trees. Import fails if the library is not built."""
from . import ffi
from .ffi import AnemoiError, ArityError, LengthError, NoDeviceError
from .fields import FIELDS, FIELD_NAMES, INST_2_1, INST_4_3
from .hasher import *  # noqa: F401,F403  (AnemoiBls12_381_2_1, ..., AnemoiDigest, HASHERS)
from .hasher import HASHERS, AnemoiDigest
from . import merkle

__version__ = "0.1.0"


def device_count():
    return ffi.lib.anemoi_b200_device_count()


def pool_reserve(device, nbytes):
    """Grow the library's device-memory pool (buffers of the host-pointer calls) ahead of the first big call."""
    ffi.check(ffi.lib.anemoi_b200_pool_reserve(int(device), int(nbytes)))


def pool_trim(device, keep_bytes=0):
    """Hand the pool's cached device memory back to the driver, keeping at most keep_bytes."""
    ffi.check(ffi.lib.anemoi_b200_pool_trim(int(device), int(keep_bytes)))
