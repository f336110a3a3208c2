#!/bin/bash
OUT=gpurun_out/run10
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -25
for b in 0 1; do echo "== HDT_BEAMS=$b"; HDT_BEAMS=$b AB_CHECK=$b timeout 900 python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'; done
echo "== launches"
HDT_BENCH_FOOTPRINT=13 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --footprint-log2 13 > $OUT/bench_ncu.log 2>&1
grep -c trace $OUT/launches.csv
