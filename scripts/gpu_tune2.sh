#!/bin/bash
# parity tests on the default build, bench (search only) on the default build and on every variant build
TAG=${1:-tune2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
line() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],1), d['search'])" || tail -3 ${1%.json}.err; }
timeout 300 python bench.py --steps 3 --warmup 2 --no-extras > $OUT/bench_default.json 2> $OUT/bench_default.err; line $OUT/bench_default.json default
for so in fuxi_planner_b200/libfuxi_b200_*.so; do
  tag=$(basename $so .so)
  FUXI_B200_SO=$PWD/$so timeout 300 python bench.py --steps 3 --warmup 2 --no-extras > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  line $OUT/bench_$tag.json $tag
done
