#!/bin/bash
OUT=gpurun_out/run13
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -5
for v in v8 v9a v9b v9c v9d; do
  echo "== $v sync per-pass"; HDT_LIB=$PWD/build/libhdt_$v.so AB_CHECK=0 timeout 900 python scripts/ab_bench.py 13 16 2>&1 | grep '^{\|rror'
  echo "== $v pipelined"; HDT_LIB=$PWD/build/libhdt_$v.so timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --footprint-log2 13 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])"
done
echo "== v8 pipelined no prefetch"; HDT_LIB=$PWD/build/libhdt_v8.so timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --footprint-log2 13 --no-beam-prefetch 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])"
