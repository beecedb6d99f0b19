#!/bin/bash
# parity tests, bench at two batch sizes, optional ncu capture of the search kernel at a steady-state batch
TAG=${1:-quick}; NCU=${2:-0}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for q in 8192 32768; do
python bench.py --steps 3 --warmup 2 --no-extras --queries $q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['queries_per_gpu'], round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],1), d['search'])"
done
if [ "$NCU" = "1" ]; then
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_search_batch -s 1 -c 1 -o $OUT/prof_search \
    python bench.py --steps 1 --warmup 1 --no-extras --queries 16384 > $OUT/ncu_search.log 2>&1; echo "ncu search rc=$?"
fi
