#!/bin/bash
OUT=gpurun_out/${1:-exp3}; mkdir -p $OUT
bash scripts/gpu_tune2.sh ${1:-exp3}
FUXI_B200_SO=$PWD/fuxi_planner_b200/libfuxi_b200_clk128.so python scripts/exp_phase.py 2>&1 | tail -9 | tee -a $OUT/phase.txt
FUXI_B200_SO=$PWD/fuxi_planner_b200/libfuxi_b200_clk128.so FUXI_SLOTS=148 python scripts/exp_phase.py 2>&1 | tail -9 | tee -a $OUT/phase.txt
