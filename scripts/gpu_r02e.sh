#!/bin/bash
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg4 or pocket or plan_host or extreme or slot_reuse or random or reference_maps or dropin" > $OUT/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -5 $OUT/pytest_quick.log
timeout 300 python scripts/lat_probe.py > $OUT/lat_cluster.txt 2>&1; echo "lat rc=$?"; cat $OUT/lat_cluster.txt
FUXI_B200_CLUSTER=0 timeout 300 python scripts/lat_probe.py 120 > $OUT/lat_nocluster.txt 2>&1; echo "lat0 rc=$?"; cat $OUT/lat_nocluster.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 600 python bench.py --config cfg5 --steps 2 --warmup 1 > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err; echo "cfg5 rc=$?"; head -c 1300 $OUT/bench_cfg5.json; echo
ls -la $OUT
