#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the bench line, the ncu launch list and ncu --set full captures.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
# launch list of the same command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
# full captures: search (one launch of the bench's own batch: 8192 queries), projection, inflation, EDT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_search_batch -s 1 -c 1 -o $OUT/prof_search \
    python bench.py --steps 1 --warmup 1 --no-extras > $OUT/ncu_search.log 2>&1; echo "ncu search rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_project|k_inflate|k_edt' -c 12 -o $OUT/prof_map \
    python scripts/map_kernels.py > $OUT/ncu_map.log 2>&1; echo "ncu map rc=$?"
ls -la $OUT
