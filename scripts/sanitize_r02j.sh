#!/bin/bash
TAG=${1:-r02j_sanitize}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONDONTWRITEBYTECODE=1
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $OUT/$tool.log python scripts/sanitize_r02j.py > $OUT/$tool.out 2>&1; echo "$tool rc=$?"
  grep "ERROR SUMMARY\|RACECHECK SUMMARY" $OUT/$tool.log | sort | uniq -c | head -3; tail -n 2 $OUT/$tool.out
done
timeout 300 python scripts/fuzz_map.py 60 2024 2>&1 | tail -3
timeout 400 python scripts/fuzz_search.py 90 2024 2>&1 | tail -3
