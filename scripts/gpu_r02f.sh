#!/bin/bash
OUT=gpurun_out/r02f
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_search_cluster -s 2 -c 1 -o $OUT/prof_cluster python scripts/lat_one.py > $OUT/ncu_cluster.log 2>&1; echo "ncu cluster rc=$?"
FUXI_B200_CLUSTER=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_search_batch -s 2 -c 1 -o $OUT/prof_lat python scripts/lat_one.py > $OUT/ncu_lat.log 2>&1; echo "ncu lat rc=$?"
tail -2 $OUT/ncu_cluster.log $OUT/ncu_lat.log
ls -la $OUT
