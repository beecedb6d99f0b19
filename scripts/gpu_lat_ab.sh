#!/bin/bash
# single-query latency on the 4096^2 grid, first-bound table off / on, cluster and one-CTA forms; + search parity tests
OUT=gpurun_out/${1:-latab}; mkdir -p $OUT
for calib in 0 1; do for cl in 1 0; do
  echo "== FUXI_B200_CALIB=$calib FUXI_B200_CLUSTER=$cl"
  FUXI_B200_CALIB=$calib FUXI_B200_CLUSTER=$cl timeout 600 python scripts/lat_probe.py 220 2>&1 | tee $OUT/lat_calib${calib}_cl${cl}.txt
done; done
timeout 1500 python -m pytest tests -m gpu -x -q -k "search or jps1 or cfg4 or cfg5 or latency or replan" > $OUT/pytest_search.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_search.log
