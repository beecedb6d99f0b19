#!/bin/bash
# ncu --set full of the EDT fast-path kernel (2 % fill launch); optional build variant tag
OUT=gpurun_out/${1:-edtprof}; mkdir -p $OUT
if [ -n "$2" ]; then export FUXI_B200_SO=$PWD/fuxi_planner_b200/libfuxi_b200_$2.so; fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_edt_(strips|bits)" -s 9 -c 1 -o $OUT/edt_main_$2 python scripts/time_edt.py > $OUT/prof.log 2>&1
tail -3 $OUT/prof.log
