#!/bin/bash
# compute-sanitizer over the batched search (search.cu, band.cu), the assembly / post-processing kernels and the map kernels
TAG=${1:-sanitize2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONDONTWRITEBYTECODE=1
FUXI_B200_SMALL=0 timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_search.log \
    python -m pytest tests/test_gpu_parity.py -x -q -k "(edge_cases and batched) or oob or (small_maps and 120 and batched) or pockets" > $OUT/memcheck_search.out 2>&1; echo "memcheck search rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_assemble.log \
    python -m pytest tests/test_gpu_assemble.py -x -q > $OUT/memcheck_assemble.out 2>&1; echo "memcheck assemble rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_map.log \
    python -m pytest tests/test_gpu_parity.py -x -q -k "(project and not large and not 1048576) or (inflate and not 2048) or (edt_exact and 256)" > $OUT/memcheck_map.out 2>&1; echo "memcheck map rc=$?"
FUXI_B200_SMALL=0 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_search.log \
    python -m pytest tests/test_gpu_parity.py -x -q -k "edge_cases and batched" > $OUT/racecheck_search.out 2>&1; echo "racecheck search rc=$?"
for f in $OUT/*.log; do echo "== $f"; grep "ERROR SUMMARY\|RACECHECK SUMMARY" $f | sort | uniq -c | head -5; done
for f in $OUT/*.out; do echo "== $f"; tail -n 1 $f; done
