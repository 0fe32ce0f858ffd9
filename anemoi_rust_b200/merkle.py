"""Jive Merkle trees on device, single- and multi-GPU (one process per GPU, torch.distributed).

The reference has no tree code; the node function is its Jive compression (arity 2: Anemoi-2-1
`compress`, src/<field>/anemoi_2_1/hasher.rs:96-103; arity 4: Anemoi-4-3 `compress_k(.,4)`,
anemoi_4_3/hasher.rs:162-179), iterated level by level, left to right.

Sharding: rank g holds leaves [g*n/G, (g+1)*n/G), reduces its slice while it still
consists of whole sub-trees (no communication), then ONE all-gather of the partial roots (<= a few
hundred bytes) and every rank finishes the top levels redundantly. The whole sharded build, NCCL
all-gather included, is one C-ABI call (anemoi_b200_merkle_root_sharded_dev) that a Rust host reaches the
same way; PyTorch is only plumbing here (device buffers, streams, the bootstrap of the communicator)."""
import ctypes

from . import ffi

_lib = ffi.lib


def plan(n_leaves, arity, world_size):
    """Host logic of the sharded build. Returns (slice_len, local_levels, roots_per_rank, top_levels)."""
    if n_leaves < 1 or world_size < 1:
        raise ffi.LengthError(ffi.ERR_LENGTH, "empty tree")
    height, m = 0, n_leaves
    while m > 1:
        if m % arity:
            raise ffi.LengthError(ffi.ERR_LENGTH, "n_leaves must be a power of the arity")
        m //= arity
        height += 1
    if n_leaves % world_size:
        raise ffi.LengthError(ffi.ERR_LENGTH, "n_leaves must be divisible by the number of ranks")
    slice_len = n_leaves // world_size
    local_levels, m = 0, slice_len
    while m > 1 and m % arity == 0:
        m //= arity
        local_levels += 1
    roots_per_rank = m
    total = roots_per_rank * world_size
    # the gathered partial roots must themselves form a complete tree
    top_levels, t = 0, total
    while t > 1:
        if t % arity:
            raise ffi.LengthError(ffi.ERR_LENGTH, "world size does not split this tree into whole sub-trees")
        t //= arity
        top_levels += 1
    assert local_levels + top_levels == height
    return slice_len, local_levels, roots_per_rank, top_levels


def merkle_reduce(H, leaves, levels, scratch=None, out=None):
    """Reduce `levels` levels on the current device. leaves: CUDA int64/uint64 tensor (n, N64)."""
    import torch

    f, arity = H.FIELD, H.STATE_WIDTH
    n = leaves.numel() // f.n64
    if n == 0 or n % (arity ** levels):
        raise ffi.LengthError(ffi.ERR_LENGTH, "n_leaves is not a multiple of arity^levels")
    n_out = n // (arity ** levels)
    if out is None:
        out = torch.empty((n_out, f.n64), dtype=leaves.dtype, device=leaves.device)
    if scratch is None:
        scratch = torch.empty((_lib.anemoi_b200_merkle_scratch_felts(arity, n), f.n64), dtype=leaves.dtype,
                              device=leaves.device)
    with torch.cuda.device(leaves.device):  # the _dev entry points launch on the CURRENT device
        stream = ctypes.c_void_p(torch.cuda.current_stream(leaves.device).cuda_stream)
        ffi.check(_lib.anemoi_b200_merkle_reduce_dev(f.id, H.INST, arity, ctypes.c_void_p(leaves.data_ptr()), n, levels,
                                                     ctypes.c_void_p(scratch.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                     stream))
    return out


def merkle_root_device(H, leaves, scratch=None):
    """Root of a device-resident tree on one GPU."""
    f, arity = H.FIELD, H.STATE_WIDTH
    n = leaves.numel() // f.n64
    _, levels, roots, _ = plan(n, arity, 1)
    assert roots == 1
    return merkle_reduce(H, leaves, levels, scratch=scratch)


_COMMS = {}   # id(process group) -> ncclComm_t created through the C ABI (kept for the life of the process)


def nccl_comm(group=None):
    """The library-side NCCL communicator of a torch.distributed process group (one rank per GPU). torch is only
    the bootstrap here: rank 0 draws an NCCL unique id through the C ABI, the 128 bytes are broadcast over the
    process group, and every rank joins with anemoi_b200_comm_init_rank -- exactly what a Rust host would do with
    its own transport. The communicator then belongs to libanemoi_b200.so; the all-gather of the sharded Merkle
    build is issued by the library on the caller's stream."""
    import torch
    import torch.distributed as dist

    key = id(group) if group is not None else 0
    if key in _COMMS:
        return _COMMS[key]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (ctypes.c_uint8 * 128)()
        ffi.check(_lib.anemoi_b200_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
        ident = torch.tensor(list(buf), dtype=torch.uint8)
    ident = ident.to(dev)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(ident, src=src, group=group)
    raw = bytes(ident.cpu().tolist())
    comm = ctypes.c_void_p()
    ffi.check(_lib.anemoi_b200_comm_init_rank(ctypes.c_char_p(raw), world, rank, ctypes.byref(comm)))
    n, r = ctypes.c_int(), ctypes.c_int()
    ffi.check(_lib.anemoi_b200_comm_info(comm, ctypes.byref(n), ctypes.byref(r)))
    assert (n.value, r.value) == (world, rank)
    _COMMS[key] = comm
    return comm


def merkle_root_distributed(H, local_leaves, group=None, reduce_fn=None, scratch=None):
    """Sharded root: every rank passes its contiguous slice (a CUDA tensor); returns the (identical) root on every
    rank. Product path: ONE C-ABI call, anemoi_b200_merkle_root_sharded_dev (per-rank sub-tree, one ncclAllGather
    of the <= 2 partial roots per rank issued by the library, top levels on every rank), on torch's current stream.
    `reduce_fn(H, tensor, levels)` exists only so that the host logic (plan -> reduce -> all-gather -> top levels)
    can be exercised on CPU/gloo in tests with the node function injected from the oracle; then the gather goes
    through torch.distributed."""
    import torch
    import torch.distributed as dist

    f, arity = H.FIELD, H.STATE_WIDTH
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_local = local_leaves.numel() // f.n64
    _, local_levels, roots_per_rank, top_levels = plan(n_local * world, arity, world)
    if reduce_fn is None:
        if not local_leaves.is_cuda or not local_leaves.is_contiguous():
            raise ValueError("local_leaves must be a contiguous CUDA tensor")
        with torch.cuda.device(local_leaves.device):
            comm = nccl_comm(group) if world > 1 else None
            root = torch.empty((1, f.n64), dtype=local_leaves.dtype, device=local_leaves.device)
            if scratch is not None:
                need = _lib.anemoi_b200_merkle_sharded_scratch_felts(arity, n_local, world)
                if scratch.numel() < need * f.n64 or scratch.device != local_leaves.device:
                    raise ValueError("scratch must hold anemoi_b200_merkle_sharded_scratch_felts() elements on the same device")
            stream = ctypes.c_void_p(torch.cuda.current_stream(local_leaves.device).cuda_stream)
            ffi.check(_lib.anemoi_b200_merkle_root_sharded_dev(
                f.id, H.INST, arity, ctypes.c_void_p(local_leaves.data_ptr()), n_local, comm,
                ctypes.c_void_p(scratch.data_ptr()) if scratch is not None else None,
                ctypes.c_void_p(root.data_ptr()), stream))
        return root
    part = reduce_fn(H, local_leaves, local_levels)
    if world == 1:
        return part if top_levels == 0 else reduce_fn(H, part, top_levels)
    gathered = torch.empty((world * roots_per_rank, f.n64), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(gathered, part.contiguous(), group=group)
    return reduce_fn(H, gathered, top_levels)
