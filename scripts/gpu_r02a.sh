#!/bin/bash
# Round-2 GPU call A: parity tests (new cfg4 / cfg5 / inline-golden / compact-path tests), smoke, bench (bidirectional vs
# unidirectional exact pass, CTA shape variants), the other BASELINE configs, ncu launch list + full capture.
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -30 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 2500 $OUT/bench.json; echo
FUXI_B200_BIDIR=0 timeout 300 python bench.py --no-extras > $OUT/bench_unidir.json 2> $OUT/bench_unidir.err; echo "unidir rc=$?"; head -c 1800 $OUT/bench_unidir.json; echo
for v in t128 t192 t512; do
  if [ -f fuxi_planner_b200/libfuxi_b200_$v.so ]; then
    FUXI_B200_SO=$PWD/fuxi_planner_b200/libfuxi_b200_$v.so timeout 300 python bench.py --no-extras > $OUT/bench_$v.json 2> $OUT/bench_$v.err; echo "$v rc=$?"; head -c 700 $OUT/bench_$v.json; echo
  fi
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
timeout 300 python bench.py --config cfg3 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; echo "cfg3 rc=$?"; head -c 1500 $OUT/bench_cfg3.json; echo
timeout 300 python bench.py --config cfg2 > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err; echo "cfg2 rc=$?"; head -c 1500 $OUT/bench_cfg2.json; echo
timeout 600 python bench.py --config cfg5 --steps 2 --warmup 1 > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err; echo "cfg5 rc=$?"; head -c 2500 $OUT/bench_cfg5.json; echo
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_search_batch -s 1 -c 1 -o $OUT/prof_search \
    python bench.py --steps 1 --warmup 1 --no-extras > $OUT/ncu_search.log 2>&1; echo "ncu search rc=$?"
ls -la $OUT
