"""Host logic of the sharded Merkle build on CPU: the partition plan for every (arity, world size) of
SURVEY.md 8(e), and the N > 1 path (plan -> local reduce -> one all-gather -> top levels) run as two
gloo ranks. The node function is injected from the oracle here (tests only); the product path always
hashes on the GPU."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from anemoi_rust_b200 import merkle
from anemoi_rust_b200.ffi import LengthError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_table():
    # arity 2, n = 2^24: every G gives one root per GPU and log2 G top levels
    for g, top in ((1, 0), (2, 1), (4, 2), (8, 3)):
        sl, local, roots, t = merkle.plan(2 ** 24, 2, g)
        assert (sl, local, roots, t) == (2 ** 24 // g, 24 - top, 1, top)
    # arity 4, n = 4^13: G = 4 -> 1 root each; G = 2 and 8 -> 2 partial roots per GPU
    assert merkle.plan(4 ** 13, 4, 1) == (4 ** 13, 13, 1, 0)
    assert merkle.plan(4 ** 13, 4, 4) == (4 ** 12, 12, 1, 1)
    assert merkle.plan(4 ** 13, 4, 2) == (2 * 4 ** 12, 12, 2, 1)
    assert merkle.plan(4 ** 13, 4, 8) == (2 * 4 ** 11, 11, 2, 2)
    assert merkle.plan(1, 4, 1) == (1, 0, 1, 0)
    with pytest.raises(LengthError):
        merkle.plan(48, 4, 1)          # not a power of the arity
    with pytest.raises(LengthError):
        merkle.plan(4 ** 3, 4, 3)      # not divisible by the world size
    with pytest.raises(LengthError):
        merkle.plan(0, 2, 1)


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import anemoi_rust_b200 as A
from anemoi_rust_b200 import merkle
from oracle import c_oracle as C

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
H = A.AnemoiPallas_4_3
f = H.FIELD
total = 4 ** 5
leaves = f.random_mont(total, 77)
local = leaves[rank * total // world:(rank + 1) * total // world]

def oracle_reduce(H, t, levels, **kw):
    a = t.numpy().view(np.uint64).reshape(-1, f.n64)
    n_out = a.shape[0] // (4 ** levels)
    if levels == 0:
        return t
    parts = [C.merkle_root(5, 1, 4, c) for c in a.reshape(n_out, -1, f.n64)]
    return torch.from_numpy(np.concatenate(parts).view(np.int64))

root = merkle.merkle_root_distributed(H, torch.from_numpy(local.view(np.int64).copy()), reduce_fn=oracle_reduce)
exp = C.merkle_root(5, 1, 4, leaves)
assert np.array_equal(root.numpy().view(np.uint64).reshape(1, -1), exp), "rank %%d root mismatch" %% rank
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_merkle(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("ok") == 2
