#!/bin/bash
# new-row parity: assembly / post-processing / fused replan tests, then the whole GPU suite if they pass
TAG=${1:-asm1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_assemble.py -x -q > $OUT/pytest_asm.log 2>&1; echo "asm rc=$?"; tail -25 $OUT/pytest_asm.log
