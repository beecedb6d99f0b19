#!/bin/bash
# experiment: (1) latency under load vs pages in use (ubench_mem), (2) search throughput vs concurrent slots
OUT=gpurun_out/exp1; mkdir -p $OUT
./scripts/ubench/ubench_mem 888 64 > $OUT/ubench_888.txt 2>&1; cat $OUT/ubench_888.txt
./scripts/ubench/ubench_mem 1776 32 > $OUT/ubench_1776.txt 2>&1; cat $OUT/ubench_1776.txt
export FUXI_B200_SO=$PWD/fuxi_planner_b200/libfuxi_b200_t128b8_cg.so
for sl in 148 296 592 888 1184; do
FUXI_SLOTS=$sl python bench.py --steps 2 --warmup 1 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slots $sl', round(d['value']), round(d['ms_per_step'],1), round(d['search']['nodes_per_s']/1e9,2))"
done
