#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (latency forms of the search incl. the cluster kernel, path forms,
# node-cloud kernel, EDT fix-list append): memcheck + racecheck + synccheck on scripts/sanitize_r02.py
TAG=${1:-r02_sanitize}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONDONTWRITEBYTECODE=1
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $OUT/$tool.log python scripts/sanitize_r02.py > $OUT/$tool.out 2>&1; echo "$tool rc=$?"
  grep "ERROR SUMMARY\|RACECHECK SUMMARY" $OUT/$tool.log | sort | uniq -c | head -3; tail -n 2 $OUT/$tool.out
done
