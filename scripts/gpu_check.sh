#!/bin/bash
# full GPU test suite + single-query latency probe + smoke (after a batch of changes)
OUT=gpurun_out/${1:-check}; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python scripts/lat_probe.py 220 2>&1 | tee $OUT/lat.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
