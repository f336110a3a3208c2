#!/bin/bash
OUT=gpurun_out/run3
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -40; echo "pytest rc=${PIPESTATUS[0]}"
echo "== golden"; timeout 300 python tests/golden/make_golden.py $OUT/golden 2>&1 | grep "pose\|Error\|error\|assert" | tail -20
echo "== bench ref-cuda F=13"; timeout 600 python bench.py --impl reference-cuda --steps 32 --warmup 8 --footprint-log2 13 2> $OUT/bench_refcuda_f13.err | grep '^{' > $OUT/bench_refcuda_f13.json; cat $OUT/bench_refcuda_f13.json
echo "== bench ref-cuda F=14"; timeout 600 python bench.py --impl reference-cuda --steps 32 --warmup 8 --footprint-log2 14 2> $OUT/bench_refcuda_f14.err | grep '^{' > $OUT/bench_refcuda_f14.json; cat $OUT/bench_refcuda_f14.json
echo "== bench ours F=14"; timeout 600 python bench.py --steps 32 --warmup 8 --footprint-log2 14 > $OUT/bench_ours_f14.json 2> $OUT/bench_ours_f14.err; cat $OUT/bench_ours_f14.json; tail -3 $OUT/bench_ours_f14.err
echo "== bench basic dag ours/ref F=13"; timeout 600 python bench.py --steps 32 --warmup 8 --footprint-log2 13 --dag basic --no-cpu-baseline 2>/dev/null | grep '^{' > $OUT/bench_ours_basic_f13.json; cat $OUT/bench_ours_basic_f13.json
timeout 600 python bench.py --impl reference-cuda --steps 32 --warmup 8 --footprint-log2 13 --dag basic 2>/dev/null | grep '^{' > $OUT/bench_refcuda_basic_f13.json; cat $OUT/bench_refcuda_basic_f13.json
echo "== ncu full (reference)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_ -s 20 -c 3 -o $OUT/prof_ref python bench.py --impl reference-cuda --steps 4 --warmup 2 --footprint-log2 13 --poses 8 > /dev/null 2>&1; ls -la $OUT | grep ncu
echo done
