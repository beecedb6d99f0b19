#!/bin/bash
# time_edt.py over every EDT build variant present (scripts/build_variants.sh edt_<R>_<inline>_<rows>_<words>)
OUT=gpurun_out/${1:-edtvar}; mkdir -p $OUT
for so in fuxi_planner_b200/libfuxi_b200_edt_*.so; do
  tag=$(basename $so .so); tag=${tag#libfuxi_b200_}
  echo "== $tag"
  FUXI_B200_SO=$PWD/$so timeout 300 python scripts/time_edt.py 2>&1 | grep -v "^exact" | tee $OUT/$tag.txt
  FUXI_B200_SO=$PWD/$so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_edt --csv --log-file $OUT/$tag.csv python scripts/time_edt.py > /dev/null 2>&1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$OUT/$tag.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split('(')[0]].append(float(r[-1].replace(',','')))
for k in ("k_edt_pack","k_edt_bits","k_edt_fix"): 
    v=agg[k]; print(k, "2%:", round(sum(v[7:17])/10000,1), "20%:", round(sum(v[17:27])/10000,1), "0.5%:", round(sum(v[27:37])/10000,1))
PY
done
