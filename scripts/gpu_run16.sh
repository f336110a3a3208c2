#!/bin/bash
OUT=gpurun_out/run16
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -5
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench default"
timeout 1500 python bench.py 2>$OUT/bench.err | tail -1 > $OUT/bench_f14.json; cat $OUT/bench_f14.json; tail -3 $OUT/bench.err
echo "== bench reference-cuda f14"
timeout 900 python bench.py --impl reference-cuda 2>/dev/null | tail -1 > $OUT/bench_refcuda_f14.json; cat $OUT/bench_refcuda_f14.json
echo "== launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 120 --csv --log-file $OUT/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --frames-in-flight 1 > $OUT/bench_ncu.log 2>&1
echo "== ncu full"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'trace_|beam_|setup_' -s 230 -c 14 -o $OUT/prof_v9 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --frames-in-flight 1 > $OUT/bench_ncu2.log 2>&1
ls -la $OUT
