"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/fuxi_b200.h
declares, the product package never touches the oracle, and it fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    import fuxi_planner_b200 as fx
    return fx


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fuxi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(built):
    lib = ctypes.CDLL(built.SO_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 16
    for s in declared:
        assert hasattr(lib, s), "libfuxi_b200.so does not export %s" % s
    from fuxi_planner_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding and header disagree"


def test_version_and_null_context(built):
    lib = built.load()
    assert lib.fx_version() >= 1
    assert lib.fx_launch_count(None) == 0
    assert lib.fx_destroy(None) == 0


def test_no_cpu_fallback_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(built.FuxiError) as ei:
        built.Context(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
    import numpy as np
    with pytest.raises(built.FuxiError):
        built.jps1.method(np.zeros((4, 4)), (0, 0), (3, 3), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fuxi_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "fuxi_oracle" not in txt, f
                assert "/root/reference" not in txt, f


def test_sass_is_sm100a(built):
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", built.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
