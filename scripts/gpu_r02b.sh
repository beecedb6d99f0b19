#!/bin/bash
# Round-2 GPU call B: parity tests (latency form, partition-form projection, EDT R = 12, cloud node product, jump points),
# bench lines of all configs, map-kernel A/B, ncu captures.
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 700 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 1800 $OUT/bench.json; echo; tail -3 $OUT/bench.err
python - <<'PY' $OUT/bench.json
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k,v in d.get("kernels",{}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms","frac")})
    print("latency", json.dumps(d.get("latency")))
    print("cpu_baseline", json.dumps(d.get("cpu_baseline")))
    print("other", json.dumps(d.get("other_configs"))[:1500])
except Exception as e: print("parse failed", e)
PY
FUXI_B200_PROJ_PART=0 timeout 300 python scripts/time_map_kernels.py > $OUT/map_red_form.txt 2>&1; echo "map A rc=$?"; cat $OUT/map_red_form.txt
timeout 300 python scripts/time_map_kernels.py > $OUT/map_part_form.txt 2>&1; echo "map B rc=$?"; cat $OUT/map_part_form.txt
timeout 300 python bench.py --config cfg3 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; echo "cfg3 rc=$?"; head -c 1500 $OUT/bench_cfg3.json; echo; tail -2 $OUT/bench_cfg3.err
timeout 600 python bench.py --config cfg5 --steps 2 --warmup 1 > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err; echo "cfg5 rc=$?"; head -c 3000 $OUT/bench_cfg5.json; echo; tail -2 $OUT/bench_cfg5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_search_batch|k_band_bound' -s 2 -c 2 -o $OUT/prof_search \
    python bench.py --steps 1 --warmup 1 --no-extras > $OUT/ncu_search.log 2>&1; echo "ncu search rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_project|k_edt|k_bits' -c 24 -o $OUT/prof_map \
    python scripts/map_kernels.py > $OUT/ncu_map.log 2>&1; echo "ncu map rc=$?"
ls -la $OUT
