#!/bin/bash
OUT=gpurun_out/run17
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== edit scenario"; timeout 900 python tests/edit_scenario.py d13 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -12
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -40
echo "== ncu full, steady state"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'trace_|beam_|setup_' -s 345 -c 21 -o $OUT/prof_v9 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --frames-in-flight 1 > $OUT/bench_ncu2.log 2>&1
echo "== launches, steady state"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 395 -c 112 --csv --log-file $OUT/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --frames-in-flight 1 > $OUT/bench_ncu.log 2>&1
ls -la $OUT
