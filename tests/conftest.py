import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the built libraries are git-ignored: on a fresh checkout build them once (nvcc cross-compiles on CPU)
    lib = os.path.join(ROOT, "anemoi_rust_b200", "libanemoi_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)
