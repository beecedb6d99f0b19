"""CPU ORACLE -- test infrastructure only, never product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``fuxi_planner_b200`` must never do so
(tests/test_boundary.py greps for it).

Contents
--------
* ``fuxi_oracle.c``  plain-C restatement of scripts/jps1.py (JPS), a Dijkstra second oracle,
  the two inflation stencils and an exact EDT; built by ``oracle.build()`` (gcc, OpenMP).
* ``hostref.py``     numpy restatements of the planner / cloud-node inline blocks
  (global_planner_st.py:226-275, global_planner_ccst.py:411-464, plc_point2_st.py:112-148,244-256).
* ``refload.py``     loader for the *unmodified* reference (only where /root/reference exists,
  i.e. the build container) -- used by tests/golden/make_golden.py and by the pin tests.

Parity pin: costs and jump-point paths in tests/golden/jps1_golden.json were produced by the
unmodified reference; tests/test_oracle.py checks the C restatement against all of them.
Projection (octomap_server / ccmapping, un-vendored, unpinned) and EDT (not in the reference)
are "parity unpinned": their semantics are defined in DESIGN.md and checked against numpy/scipy.
"""
from .capi import (build, lib_path, jps, jps_batch, sssp_field, sssp_cost, sssp_batch,
                   inflate, edt, num_threads, cloud_filter, jump)
from . import hostref

__all__ = ["build", "lib_path", "jps", "jps_batch", "sssp_field", "sssp_cost", "sssp_batch",
           "inflate", "edt", "num_threads", "cloud_filter", "jump", "hostref"]
