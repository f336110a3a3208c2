#!/bin/bash
OUT=gpurun_out/run15
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "leaked\|^\s*$\|took\|Creating\|reallocating\|pool size\|page table size\|bloom\|color nodes\|leaves$" | tail -5
S='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step","wall_ms_per_step","gpu_launches")}, d["e2e"]["ms_per_step"])'
for fl in "--frames-in-flight 1 --no-beam-prefetch" "--frames-in-flight 1" "--frames-in-flight 2 --no-beam-prefetch" "--frames-in-flight 2"; do
  echo "== $fl"; timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --footprint-log2 13 $fl 2>&1 | tail -1 | python -c "$S"
done
echo "== beams off, 2 in flight"; HDT_BEAMS=0 timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --footprint-log2 13 2>&1 | tail -1 | python -c "$S"
echo "== f14 default"; timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
echo "== f14 4K"; timeout 900 python bench.py --steps 32 --warmup 8 --no-cpu-baseline --width 3840 --height 2160 2>&1 | tail -1 | python -c "$S"
