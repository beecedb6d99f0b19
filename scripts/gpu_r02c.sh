#!/bin/bash
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 300 python scripts/lat_probe.py > $OUT/lat_default.txt 2>&1; echo "lat rc=$?"; cat $OUT/lat_default.txt
for v in w1024 w256 sq0; do
  if [ -f fuxi_planner_b200/libfuxi_b200_$v.so ]; then
    FUXI_B200_SO=$PWD/fuxi_planner_b200/libfuxi_b200_$v.so timeout 300 python scripts/lat_probe.py > $OUT/lat_$v.txt 2>&1; echo "$v rc=$?"; cat $OUT/lat_$v.txt
  fi
done
timeout 700 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 1700 $OUT/bench.json; echo; tail -3 $OUT/bench.err
python - <<'PY' $OUT/bench.json
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"], "band", d["roofline"].get("band_kernel_ms"), "search", d["roofline"]["kernel_ms"])
    print("latency", json.dumps(d.get("latency")))
except Exception as e: print("parse failed", e)
PY
timeout 600 python bench.py --config cfg5 --steps 2 --warmup 1 > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err; echo "cfg5 rc=$?"; head -c 1200 $OUT/bench_cfg5.json; echo
ls -la $OUT
