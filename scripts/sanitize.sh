#!/bin/bash
# compute-sanitizer over the newer kernels (cloud conditioning, small-map search, formats, EDT passes): memcheck everywhere,
# racecheck on the shared-memory search.  Usage on the GPU box: bash scripts/sanitize.sh [tag]
TAG=${1:-sanitize}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONDONTWRITEBYTECODE=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_cloud.log \
    python -m pytest tests/test_gpu_cloud.py tests/test_gpu_formats.py -x -q -k "not 150000 and not 60000" > $OUT/memcheck_cloud.out 2>&1; echo "memcheck cloud/formats rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_small.log \
    python -m pytest tests/test_gpu_parity.py -x -q -k "small_maps or edge_cases or oob or dropin or separable or cfg1" > $OUT/memcheck_small.out 2>&1; echo "memcheck small/edt rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_small.log \
    python -m pytest tests/test_gpu_parity.py -x -q -k "edge_cases and shared" > $OUT/racecheck_small.out 2>&1; echo "racecheck small rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_cloud.log \
    python -m pytest tests/test_gpu_cloud.py -x -q -k "scene and 5000 or golden or parameters" > $OUT/racecheck_cloud.out 2>&1; echo "racecheck cloud rc=$?"
for f in $OUT/*.log; do echo "== $f"; grep -c "ERROR SUMMARY" $f; grep "ERROR SUMMARY\|RACECHECK SUMMARY" $f | sort | uniq -c | head -5; done
for f in $OUT/*.out; do tail -n 2 $f; done
