#!/bin/bash
OUT=gpurun_out/run12
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -25
echo "== beams off"; HDT_BEAMS=0 AB_CHECK=0 timeout 900 python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'
for cap in 16 24 32 48; do echo "== cap $cap"; HDT_BEAM_MAX_VISITS=$cap AB_CHECK=$((cap==32)) timeout 900 python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'; done
echo "== bench"
timeout 900 python bench.py --steps 32 --warmup 8 --no-cpu-baseline --footprint-log2 13 2>$OUT/bench.err | tail -1 > $OUT/bench_f13.json; cat $OUT/bench_f13.json
echo "== launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --footprint-log2 13 > $OUT/bench_ncu.log 2>&1
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'trace_paths|trace_shadows|setup_' -s 160 -c 8 -o $OUT/prof_v8 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --footprint-log2 13 > $OUT/bench_ncu2.log 2>&1
ls -la $OUT
