#!/bin/bash
OUT=gpurun_out/exp2; mkdir -p $OUT
for so in clk128 clk256; do for sl in 148 0; do
FUXI_B200_SO=$PWD/fuxi_planner_b200/libfuxi_b200_$so.so FUXI_SLOTS=$sl python scripts/exp_phase.py 2>&1 | tail -9 | tee -a $OUT/phase.txt
done; done
