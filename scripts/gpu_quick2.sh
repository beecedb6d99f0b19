#!/bin/bash
# parity tests + search-only bench + kernel launch list (ncu time metric only)
TAG=${1:-q2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
python bench.py --steps 3 --warmup 2 --no-extras 2>$OUT/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],1), d['search'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_launches.log 2>&1
grep -E "k_search|k_band|k_build|k_order" $OUT/launches.csv | awk -F'","' '{print $5, $NF}' | tail -4
