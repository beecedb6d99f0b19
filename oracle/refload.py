"""Loader for the UNMODIFIED reference (scripts/jps1.py) -- build container only.

CPU ORACLE -- test infrastructure only.  /root/reference does not exist on the GPU box, so nothing
that runs there may depend on this module; it is used by tests/golden/make_golden.py (which writes
the committed fixtures) and by the ``reference``-marked pin tests that skip when the tree is absent.
"""
import contextlib
import importlib.util
import io
import os
import sys

REF_ROOT = os.environ.get("FUXI_REFERENCE_ROOT", "/root/reference")
_jps1 = None


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "scripts", "jps1.py"))


def jps1_module():
    """Import scripts/jps1.py by path without writing bytecode into the read-only tree."""
    global _jps1
    if _jps1 is None:
        old = sys.dont_write_bytecode
        sys.dont_write_bytecode = True
        try:
            spec = importlib.util.spec_from_file_location("_fuxi_reference_jps1",
                                                          os.path.join(REF_ROOT, "scripts", "jps1.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        finally:
            sys.dont_write_bytecode = old
        _jps1 = mod
    return _jps1


def method(matrix, start, goal, hchoice):
    """Run the reference's jps1.method; its only cost output is the print at jps1.py:207, captured here.
    Returns (path or 0, cost float or None, secs)."""
    mod = jps1_module()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        path, secs = mod.method(matrix, start, goal, hchoice)
    if path == 0 and not isinstance(path, list):
        return 0, None, secs
    return path, float(buf.getvalue().strip().splitlines()[-1]), secs


def map_files():
    d = os.path.join(REF_ROOT, "maps")
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".png"))
