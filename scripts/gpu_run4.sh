#!/bin/bash
OUT=gpurun_out/run5
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
for v in $@; do HDT_LIB=$PWD/build/libhdt_$v.so python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'; done
