#!/bin/bash
# bench.py --no-extras for the default build and every variant .so present; prints value / kernel ms / parity
OUT=gpurun_out/${1:-ab}; mkdir -p $OUT
for so in fuxi_planner_b200/libfuxi_b200.so fuxi_planner_b200/libfuxi_b200_*.so; do
  [ -f $so ] || continue
  tag=$(basename $so .so)
  FUXI_B200_SO=$PWD/$so timeout 600 python bench.py --no-extras --steps 5 --warmup 3 > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
d=json.loads(open("$OUT/$tag.json").read().strip().splitlines()[-1])
print("$tag", "value %.0f"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel_ms %.2f"%d["roofline"]["kernel_ms"], "band %.2f"%d["roofline"]["band_kernel_ms"], "parity_mismatches", d.get("parity_mismatches"), "e2e %.0f"%d["e2e"]["value"])
PY
done
