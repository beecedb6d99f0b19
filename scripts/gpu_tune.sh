#!/bin/bash
# Tuning sweep on one B200: parity tests on the default build, then the bench line per build variant.
TAG=${1:-tune}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 2 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
print("default", d["value"], d["e2e"]["value"], d["search"])
for k,v in d.get("kernels",{}).items(): print(k, round(v["ms"],4), round(v["frac"],3))
print(d.get("latency")); print(d.get("cpu_baseline"))
PY
for so in fuxi_planner_b200/libfuxi_b200_*.so; do
  tag=$(basename $so .so)
  FUXI_B200_SO=$PWD/$so timeout 300 python bench.py --steps 2 --warmup 1 --no-extras > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python -c "
import json,sys
d=json.loads(open('$OUT/bench_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value']), round(d['e2e']['value']), d['search']['settled_cells_per_step_rank0'], d['search']['passes'])" || tail -3 $OUT/bench_$tag.err
done
