#!/bin/bash
# Final single-GPU round of round 2: tests, smoke, every bench line (cfg4 default, CPU arm, cfg2 / cfg3 / cfg5), launch list,
# ncu --set full of the search kernel (traffic) and of the map kernels incl. the rewritten EDT.
TAG=${1:-r02j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -2 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
for c in cfg2 cfg3 cfg5; do timeout 900 python bench.py --config $c > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "$c rc=$?"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_search_batch|k_band_bound' -s 2 -c 2 -o $OUT/prof_search \
    python bench.py --steps 1 --warmup 1 --no-extras > $OUT/ncu_search.log 2>&1; echo "ncu search rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_project|k_inflate|k_edt|k_bits' -c 16 -o $OUT/prof_map \
    python scripts/map_kernels.py > $OUT/ncu_map.log 2>&1; echo "ncu map rc=$?"
ls $OUT
