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


@pytest.fixture(scope="session")
def inline_golden():
    """a10/a11/a13/a14: outputs of the reference's own inline source lines (tests/golden/make_inline_golden.py)."""
    import base64
    import zlib
    with open(os.path.join(GOLDEN, "inline_golden.json")) as fh:
        d = json.load(fh)

    def arr(s, shape):
        return np.frombuffer(zlib.decompress(base64.b64decode(s)), dtype=np.uint8).reshape(shape).copy()

    cases = []
    for r in d["cases"]:
        c = dict(r)
        c["X"] = arr(r["X"], (r["W"], r["H"]))
        c["map_o"] = [float(v) for v in r["map_o"]]
        c["reso"] = float(r["reso"])
        c["start"] = [float(v) for v in r["start"]]
        c["goal"] = [float(v) for v in r["goal"]]
        if r["out"] is not None:
            o = dict(r["out"])
            o["grid"] = arr(o["grid"], tuple(o["shape"]))
            o["map_o"] = [float(v) for v in o["map_o"]]
            c["out"] = o
        cases.append(c)
    return cases


@pytest.fixture(scope="session")
def cfg4_golden():
    with open(os.path.join(GOLDEN, "cfg4_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def cloud_golden():
    """Outputs of the reference's own cloud-node lines (tests/golden/make_cloud_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "cloud_golden.npz"))
    cases = []
    for i in range(int(z["ncase"][0])):
        par = z["par_%d" % i]
        cases.append(dict(cam=z["cam_%d" % i], rpy=par[0:3], pos=par[3:6], local_pos=par[6:9], ang_vel=par[9:12], line_vel=par[12:15],
                          dt=float(par[15]), out=z["out_%d" % i], cen=z["cen_%d" % i], octo=z["octo_%d" % i]))
    return cases
