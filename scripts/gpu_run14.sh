#!/bin/bash
OUT=gpurun_out/run14
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
HDT_LIB=$PWD/build/libhdt_dbg.so timeout 600 python scripts/timeline.py 13 2>&1 | grep -v "^$"
