#!/bin/bash
# EDT iteration: tests, timings, per-kernel ncu times (+ optional full capture of k_edt_bits)
OUT=gpurun_out/${1:-edt2}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "edt" > $OUT/pytest_edt.log 2>&1; echo "pytest edt rc=$?"; tail -3 $OUT/pytest_edt.log
timeout 600 python scripts/time_edt.py 2>&1 | tee $OUT/time_edt.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_edt --csv --log-file $OUT/edt_launches.csv python scripts/time_edt.py > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$OUT/edt_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split('(')[0]].append(float(r[-1].replace(',','')))
for k,v in agg.items(): print(k, len(v), "last5 us:", [round(x/1000,1) for x in v[-5:]])
PY
if [ "$2" = "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_edt_bits -s 30 -c 1 -o $OUT/edt_bits python scripts/time_edt.py > /dev/null 2>&1
fi
