"""Multi-GPU paths on real devices (skipped on a single-GPU box): the single-process
anemoi_b200_merkle_root(..., n_gpus) and the one-process-per-GPU NCCL build of merkle.py, both against the
1-GPU root and the oracle."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import anemoi_rust_b200 as A

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    return A.device_count()


@pytest.mark.skipif(n_gpus() < 2, reason="needs 2 GPUs")
def test_single_process_multi_gpu_root():
    from oracle import c_oracle as C

    for H, fi, ii, h in ((A.AnemoiPallas_4_3, 5, 1, 6), (A.AnemoiBls12_377_2_1, 0, 0, 10)):
        f, ar = H.FIELD, H.STATE_WIDTH
        leaves = f.random_mont(ar ** h, 11)
        r1 = H.merkle_root(leaves, 1)
        assert np.array_equal(r1, C.merkle_root(fi, ii, ar, leaves))
        g = 2
        while g <= n_gpus():
            assert np.array_equal(H.merkle_root(leaves, g), r1), "n_gpus=%d" % g
            g *= 2


@pytest.mark.skipif(n_gpus() < 2, reason="needs 2 GPUs")
def test_single_process_multi_gpu_root_with_chunked_upload():
    """Slices big enough (>= 2^20 level-1 nodes per GPU) that every device uploads its slice in chunks and hashes the first
    level while the next chunk is in flight, then joins the all-gather from its level-1 array: same root as the
    device-resident single-GPU build (Pallas 4-3: 2 partial roots per rank; BLS12-377 2-1: 1 per rank)."""
    import torch

    from anemoi_rust_b200 import merkle

    for H, h in ((A.AnemoiPallas_4_3, 12), (A.AnemoiBls12_377_2_1, 22)):
        f, ar = H.FIELD, H.STATE_WIDTH
        leaves = f.random_mont(ar ** h, 12)
        dev = merkle.merkle_root_device(H, torch.from_numpy(leaves.view(np.int64)).cuda())
        torch.cuda.synchronize()
        exp = dev.cpu().numpy().view(np.uint64)
        del dev
        assert np.array_equal(H.merkle_root(leaves, 2), exp)


@pytest.mark.skipif(n_gpus() < 2, reason="needs 2 GPUs")
def test_single_process_multi_gpu_compress():
    H = A.AnemoiBn254_4_3
    f = H.FIELD
    x = f.random_mont(4 * 3001, 21)   # odd count: uneven slices
    one = H.compress_k_batch(x, 4)
    g = 2
    while g <= n_gpus():
        assert np.array_equal(H.compress_k_batch(x, 4, n_gpus=g), one)
        g *= 2


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import anemoi_rust_b200 as A
from anemoi_rust_b200 import merkle
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
H = A.AnemoiVesta_4_3
f = H.FIELD
total = 4 ** 7
leaves = f.random_mont(total, 123)
local = torch.from_numpy(leaves[rank * total // world:(rank + 1) * total // world].view(np.int64).copy()).cuda()
root = merkle.merkle_root_distributed(H, local)
torch.cuda.synchronize()
if rank == 0:
    full = merkle.merkle_root_device(H, torch.from_numpy(leaves.view(np.int64)).cuda())
    assert torch.equal(root.reshape(-1), full.reshape(-1)), "sharded root != single-GPU root"
    print("nccl-merkle-ok")
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(n_gpus() < 2, reason="needs 2 GPUs")
def test_nccl_sharded_root(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "nccl-merkle-ok" in out.stdout


@pytest.mark.skipif(n_gpus() < 2, reason="needs 2 GPUs")
def test_sharded_dev_entry_with_library_communicators():
    """anemoi_b200_merkle_root_sharded_dev driven the way a Rust host would: unique id from rank 0, every rank (here: one
    host thread per GPU) joins with anemoi_b200_comm_init_rank and makes ONE call; all ranks end with the 1-GPU root."""
    import ctypes
    import threading

    import torch

    from anemoi_rust_b200 import ffi, merkle
    from oracle import c_oracle as C

    lib = ffi.lib
    assert lib.anemoi_b200_nccl_version() > 0
    for H, fi, ii, h in ((A.AnemoiPallas_4_3, 5, 1, 6), (A.AnemoiBls12_381_2_1, 1, 0, 9)):
        f, ar = H.FIELD, H.STATE_WIDTH
        total = ar ** h
        leaves = f.random_mont(total, 31)
        exp = C.merkle_root(fi, ii, ar, leaves)
        world = 2
        ident = (ctypes.c_uint8 * 128)()
        ffi.check(lib.anemoi_b200_comm_unique_id(ctypes.cast(ident, ctypes.c_void_p)))
        roots, errs = [None] * world, [None] * world

        def rank_main(g):
            try:
                torch.cuda.set_device(g)
                comm = ctypes.c_void_p()
                ffi.check(lib.anemoi_b200_comm_init_rank(ctypes.cast(ident, ctypes.c_void_p), world, g, ctypes.byref(comm)))
                sl = total // world
                local = torch.from_numpy(leaves[g * sl:(g + 1) * sl].view(np.int64).copy()).to("cuda:%d" % g)
                root = torch.empty((1, f.n64), dtype=torch.int64, device="cuda:%d" % g)
                st = torch.cuda.current_stream(torch.device("cuda", g)).cuda_stream
                ffi.check(lib.anemoi_b200_merkle_root_sharded_dev(f.id, H.INST, ar, ctypes.c_void_p(local.data_ptr()), sl, comm,
                                                                  None, ctypes.c_void_p(root.data_ptr()), ctypes.c_void_p(st)))
                torch.cuda.synchronize(g)
                roots[g] = root.cpu().numpy().view(np.uint64)
                ffi.check(lib.anemoi_b200_comm_destroy(comm))
            except Exception as exc:  # surfaced below
                errs[g] = exc

        th = [threading.Thread(target=rank_main, args=(g,)) for g in range(world)]
        [t.start() for t in th]
        [t.join(timeout=300) for t in th]
        assert errs == [None] * world, errs
        for g in range(world):
            assert np.array_equal(roots[g], exp), "rank %d" % g
