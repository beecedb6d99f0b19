#!/bin/bash
for nt in 1 0; do for th in 16 8 4; do
echo "== NT=$nt host threads=$th"
FUXI_B200_NT=$nt FUXI_B200_HOST_THREADS=$th python scripts/lat_trace.py 2>&1 | grep trace | sed -n "6p;9p;36p;39p" | sed 's/fx_plan_host trace: //'
done; done
