import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the unmodified reference tree (/root/reference)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "jps1_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def maps():
    z = np.load(os.path.join(GOLDEN, "maps.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def hostfn_golden():
    z = np.load(os.path.join(GOLDEN, "hostfn_golden.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


def unpack_grid(rec):
    W, H = rec["W"], rec["H"]
    return np.unpackbits(np.array(rec["grid"], dtype=np.uint8))[: W * H].reshape(W, H)


def large_grid(rec):
    n = rec["n"]
    return (np.random.default_rng(rec["grid_seed"]).random((n, n)) < 0.2).astype(np.uint8)
