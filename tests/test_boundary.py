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


def test_header_is_plain_c_and_structs_match_the_ctypes_mirrors(tmp_path):
    """include/fuxi_b200.h must compile as C99 (it is the FFI surface), and the ctypes mirrors of its structs must have
    the C compiler's sizes and field offsets."""
    import ctypes as C
    import subprocess
    from fuxi_planner_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "probe.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "fuxi_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu\\n", sizeof(fx_replan_in), sizeof(fx_replan_out), sizeof(fx_cloud_params), offsetof(fx_cloud_params, radius));
    printf("%zu %zu %zu\\n", offsetof(fx_replan_in, origin_x), offsetof(fx_replan_out, cost_f), offsetof(fx_replan_out, origin_x));
    return 0;
}
''')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(v) for v in out]
    want = [C.sizeof(_lib.ReplanIn), C.sizeof(_lib.ReplanOut), C.sizeof(_lib.CloudParams), _lib.CloudParams.radius.offset,
            _lib.ReplanIn.origin_x.offset, _lib.ReplanOut.cost_f.offset, _lib.ReplanOut.origin_x.offset]
    assert got == want, (got, want)


def test_ctypes_argument_counts_match_the_header():
    """Every prototype in include/fuxi_b200.h has as many parameters as the argtypes list bound to it in _lib.py
    (ctypes would otherwise mis-call silently), and pointer / integer / floating parameters sit at the same positions."""
    import ctypes as C
    import re
    from fuxi_planner_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "fuxi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    lib = _lib.load()
    seen = 0
    for m in re.finditer(r"\b(?:int|int64_t|const char \*)\s*(fx_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        argtypes = getattr(lib, name).argtypes
        assert argtypes is not None, name
        assert len(argtypes) == len(plist), (name, len(argtypes), plist)
        for p, t in zip(plist, argtypes):
            is_ptr = "*" in p
            if is_ptr:
                assert t is C.c_void_p or isinstance(t, type) and issubclass(t, (C._Pointer, C.c_char_p.__class__)) or t is C.c_char_p, (name, p, t)
            elif re.search(r"\b(float|double)\b", p):
                assert t in (C.c_float, C.c_double), (name, p, t)
                assert (t is C.c_double) == bool(re.search(r"\bdouble\b", p)), (name, p, t)
            else:
                assert t in (C.c_int, C.c_int64, C.c_size_t, C.c_int32), (name, p, t)
                if re.search(r"\b(int64_t|size_t)\b", p):
                    assert t in (C.c_int64, C.c_size_t), (name, p, t)
        seen += 1
    assert seen == len(_lib.SYMBOLS), (seen, len(_lib.SYMBOLS))
