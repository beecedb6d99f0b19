#!/bin/bash
# Tuning builds of the same sources with different compile-time knobs -> fuxi_planner_b200/libfuxi_b200_<tag>.so
set -e
cd "$(dirname "$0")/.."
rm -f fuxi_planner_b200/libfuxi_b200_*.so
b() { tag=$1; shift; python fuxi_planner_b200/build.py --out=libfuxi_b200_$tag.so "$@" > /dev/null 2>&1; echo built $tag "$@"; }
for v in "$@"; do
  case $v in
    t256b3_cg) b $v -DFX_SEARCH_THREADS=256 -DFX_SEARCH_MINB=3;;
    t128b6_ca) b $v -DFX_SEARCH_THREADS=128 -DFX_SEARCH_MINB=6 -DFX_LDF=__ldca;;
    t256b4_cg) b $v -DFX_SEARCH_THREADS=256 -DFX_SEARCH_MINB=4;;
    t128b6_cg) b $v -DFX_SEARCH_THREADS=128 -DFX_SEARCH_MINB=6;;
    t64b10_cg) b $v -DFX_SEARCH_THREADS=64 -DFX_SEARCH_MINB=10;;
    lazy) b $v -DFX_EAGER_PROBE=0;;
    t128b5_cg) b $v -DFX_SEARCH_THREADS=128 -DFX_SEARCH_MINB=5;;
    rowmajor) b $v -DFX_TILED=0;;
    t128b8_cg) b $v -DFX_SEARCH_THREADS=128 -DFX_SEARCH_MINB=8;;
    t128b12_cg) b $v -DFX_SEARCH_THREADS=128 -DFX_SEARCH_MINB=12;;
    t64b12_cg) b $v -DFX_SEARCH_THREADS=64 -DFX_SEARCH_MINB=12;;
    t512b1_cg) b $v -DFX_SEARCH_THREADS=512 -DFX_SEARCH_MINB=1;;
    t128b10_cg) b $v -DFX_SEARCH_THREADS=128 -DFX_SEARCH_MINB=10;;
    t256b5_cg) b $v -DFX_SEARCH_THREADS=256 -DFX_SEARCH_MINB=5;;
    clk128) b $v -DFX_PHASE_CLOCKS -DFX_SEARCH_MINB=8;;
    edt_*) # edt_<R>_<inline>_<rows>_<words>
      IFS=_ read -r _ r inl rows words <<< "$v"; b $v -DEDT_R=$r -DEDT_INLINE_FIX=${inl}u -DEDT_BT_ROWS=$rows -DEDT_BT_WORDS=$words;;
    *) echo unknown variant $v; exit 1;;
  esac
done
