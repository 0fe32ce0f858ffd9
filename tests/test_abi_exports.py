"""The C-ABI library loads and exports every symbol include/anemoi_b200.h declares (no compute calls:
this runs without a GPU), and the introspection entries agree with the reference's constants."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "anemoi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(anemoi_b200_\w+)\s*\(", src)))


def test_header_symbols_are_exported():
    names = declared_functions()
    assert len(names) >= 40
    lib = ctypes.CDLL(os.path.join(ROOT, "anemoi_rust_b200", "libanemoi_b200.so"))
    for n in names:
        assert hasattr(lib, n), "libanemoi_b200.so does not export " + n


def test_python_binding_covers_header():
    from anemoi_rust_b200 import ffi

    assert sorted(ffi.EXPORTED) == declared_functions()


def test_introspection_matches_reference_constants():
    import json

    from anemoi_rust_b200 import ffi, FIELD_NAMES

    lib = ffi.lib
    params = json.load(open(os.path.join(ROOT, "tests", "golden", "params.json")))
    assert lib.anemoi_b200_version() == 100
    for i, name in enumerate(FIELD_NAMES):
        assert lib.anemoi_b200_field_name(i).decode() == name
        assert lib.anemoi_b200_field_limbs(i) == params[name]["n64"]
        for inst, key in ((0, "anemoi_2_1"), (1, "anemoi_4_3")):
            assert lib.anemoi_b200_num_rounds(i, inst) == params[name]["inst"][key]["rounds"]
            assert lib.anemoi_b200_state_width(inst) == params[name]["inst"][key]["width"]
            assert lib.anemoi_b200_rate_width(inst) == params[name]["inst"][key]["rate"]
    assert lib.anemoi_b200_field_limbs(7) == ffi.ERR_FIELD
    assert lib.anemoi_b200_num_rounds(0, 2) == ffi.ERR_INST
    assert b"no CPU fallback" in lib.anemoi_b200_strerror(ffi.ERR_NO_DEVICE)
    assert lib.anemoi_b200_merkle_scratch_felts(4, 4 ** 6) >= 4 ** 5 + 4 ** 4


def test_argument_errors_without_a_device():
    """Argument validation happens before any CUDA call, so it is observable on a CPU-only box; a valid
    call must then fail loudly with ERR_NO_DEVICE (there is no CPU fallback)."""
    import numpy as np

    from anemoi_rust_b200 import ffi

    lib = ffi.lib
    z = np.zeros(64, dtype=np.uint64)
    p = ctypes.c_void_p(z.ctypes.data)
    assert lib.anemoi_b200_compress(9, 0, 2, p, p, 1, 0) == ffi.ERR_FIELD
    assert lib.anemoi_b200_compress(1, 5, 2, p, p, 1, 0) == ffi.ERR_INST
    assert lib.anemoi_b200_compress(1, 0, 4, p, p, 1, 0) == ffi.ERR_ARITY     # hasher.rs:107
    assert lib.anemoi_b200_compress(1, 1, 3, p, p, 1, 0) == ffi.ERR_ARITY     # 4-3 hasher.rs:163-165
    assert lib.anemoi_b200_merkle_root(1, 0, 2, p, 6, p, 1) == ffi.ERR_LENGTH
    assert lib.anemoi_b200_merkle_root(5, 1, 2, p, 4, p, 1) == ffi.ERR_ARITY
    if lib.anemoi_b200_device_count() == 0:
        assert lib.anemoi_b200_compress(1, 0, 2, p, p, 1, 0) == ffi.ERR_NO_DEVICE
        import anemoi_rust_b200 as A

        with pytest.raises(A.NoDeviceError):
            A.AnemoiBls12_381_2_1.compress([0, 1])
        # the operational entries too: nothing pretends to have reserved device memory
        assert lib.anemoi_b200_pool_reserve(0, 1 << 20) == ffi.ERR_NO_DEVICE
        assert lib.anemoi_b200_pool_trim(0, 0) == ffi.ERR_NO_DEVICE
        with pytest.raises(A.NoDeviceError):
            A.pool_reserve(0, 1 << 20)


def test_sharded_plan_in_c_matches_python_plan():
    """The partition plan lives twice: merkle.plan (host logic, gloo-tested) and shard_plan in api.cu (what the C-ABI call
    uses). anemoi_b200_merkle_sharded_scratch_felts exposes the C side's view: local scratch + partial roots + gathered
    roots + top-level scratch, 0 when the ranks do not split the tree into whole sub-trees."""
    from anemoi_rust_b200 import ffi, merkle

    lib = ffi.lib
    for arity, total, world in ((4, 4 ** 13, 1), (4, 4 ** 13, 2), (4, 4 ** 13, 4), (4, 4 ** 13, 8), (2, 2 ** 24, 8), (2, 2 ** 10, 2),
                                (4, 4 ** 5, 2), (4, 4 ** 3, 1)):
        n_local = total // world
        _, _, roots, _ = merkle.plan(total, arity, world)
        gathered = roots * world
        exp = (lib.anemoi_b200_merkle_scratch_felts(arity, n_local) + roots + gathered +
               lib.anemoi_b200_merkle_scratch_felts(arity, gathered))
        assert lib.anemoi_b200_merkle_sharded_scratch_felts(arity, n_local, world) == exp
    assert lib.anemoi_b200_merkle_sharded_scratch_felts(4, 3 * 4 ** 3, 1) == 0      # 192 leaves: not a power of 4
    assert lib.anemoi_b200_merkle_sharded_scratch_felts(4, 4 ** 3, 3) == 0          # 3 ranks x 64 leaves
    assert lib.anemoi_b200_merkle_sharded_scratch_felts(4, 0, 2) == 0
    # argument errors of the sharded entry are reported before any CUDA / NCCL call
    import ctypes

    import numpy as np

    z = np.zeros(8, dtype=np.uint64)
    p = ctypes.c_void_p(z.ctypes.data)
    assert lib.anemoi_b200_merkle_root_sharded_dev(5, 1, 2, p, 16, None, None, p, None) == ffi.ERR_ARITY
    assert lib.anemoi_b200_merkle_root_sharded_dev(5, 1, 4, p, 0, None, None, p, None) == ffi.ERR_LENGTH
    assert lib.anemoi_b200_merkle_root_sharded_dev(5, 1, 4, None, 16, None, None, p, None) == ffi.ERR_ARG
    assert lib.anemoi_b200_merkle_root_sharded_dev(5, 1, 4, p, 48, None, None, p, None) == ffi.ERR_LENGTH
    assert lib.anemoi_b200_count_noncanonical(9, p, 1, p, 0) == ffi.ERR_FIELD
    assert lib.anemoi_b200_comm_init_rank(None, 2, 0, None) == ffi.ERR_ARG
