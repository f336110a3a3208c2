#!/bin/bash
OUT=gpurun_out/run4
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
for v in v1 v2; do HDT_LIB=$PWD/build/libhdt_$v.so python scripts/ab_bench.py 13 16 2>&1 | grep '^{'; done
for mb in 4 5 6; do HDT_LIB=$PWD/build/libhdt_v2_mb$mb.so HDT_PERSISTENT=$mb AB_CHECK=1 python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'; done
HDT_LIB=$PWD/build/libhdt_v2_mb5.so HDT_PERSISTENT=10 AB_CHECK=0 python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'
HDT_LIB=$PWD/build/libhdt_v2_mb6.so HDT_PERSISTENT=12 AB_CHECK=0 python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'
echo "== pytest (v2 default lib)"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -15
echo "== ncu v2 + persistent"; HDT_LIB=$PWD/build/libhdt_v2_mb5.so HDT_PERSISTENT=5 AB_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_paths -s 40 -c 1 -o $OUT/prof_persist python scripts/ab_bench.py 13 8 > /dev/null 2>&1
HDT_LIB=$PWD/build/libhdt_v2.so AB_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_paths -s 40 -c 1 -o $OUT/prof_v2 python scripts/ab_bench.py 13 8 > /dev/null 2>&1
ls -la $OUT
