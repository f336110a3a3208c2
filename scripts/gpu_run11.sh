#!/bin/bash
OUT=gpurun_out/run11
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -25
for cap in 16 24 32 48 1000; do echo "== cap $cap"; HDT_BEAM_MAX_VISITS=$cap AB_CHECK=0 timeout 900 python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'; done
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'trace_paths|beam_paths|seed_|trace_shadows|beam_shadows' -s 40 -c 12 -o $OUT/prof_v7 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --footprint-log2 13 > $OUT/bench_ncu.log 2>&1
ls -la $OUT
