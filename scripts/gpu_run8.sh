#!/bin/bash
OUT=gpurun_out/run9
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -15
for v in v3 v5; do HDT_LIB=$PWD/build/libhdt_$v.so AB_CHECK=0 python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'; done
