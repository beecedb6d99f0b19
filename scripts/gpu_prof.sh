#!/bin/bash
# parity tests + ncu full captures of the search kernel (one build variant) and the map kernels
TAG=${1:-prof}
SO=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
[ -n "$SO" ] && export FUXI_B200_SO=$PWD/$SO
timeout 300 python bench.py --steps 2 --warmup 1 --no-extras > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 600 $OUT/bench.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_search_batch -s 1 -c 1 -o $OUT/prof_search \
    python bench.py --steps 1 --warmup 1 --no-extras --queries 2048 > $OUT/ncu_search.log 2>&1; echo "ncu search rc=$?"
unset FUXI_B200_SO
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_project|k_inflate|k_edt|k_bits' -c 16 -o $OUT/prof_map \
    python scripts/map_kernels.py > $OUT/ncu_map.log 2>&1; echo "ncu map rc=$?"
timeout 300 python bench.py --steps 1 --warmup 1 --queries 512 --cpu-seconds 1 > $OUT/bench_extras.json 2> $OUT/bench_extras.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_extras.json").read().strip().splitlines()[-1])
for k,v in d.get("kernels",{}).items(): print(k, round(v["ms"],4), round(v["frac"],3))
PY
