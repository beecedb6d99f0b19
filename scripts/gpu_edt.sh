#!/bin/bash
OUT=gpurun_out/${1:-edt1}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "edt" > $OUT/pytest_edt.log 2>&1; echo "pytest edt rc=$?"; tail -5 $OUT/pytest_edt.log
timeout 300 python bench.py --steps 1 --warmup 1 --queries 512 --cpu-seconds 1 > $OUT/bench_extras.json 2> $OUT/bench_extras.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_extras.json").read().strip().splitlines()[-1])
for k,v in d.get("kernels",{}).items(): print(k, round(v["ms"],4), round(v["frac"],3))
PY
