#!/bin/bash
# Build library variants from compile-time switches (here: nvcc only) and time them on the GPU box (scripts/ab_bench.py).
#   here:        bash scripts/ab_variants.sh build                 # writes ab_libs/lib_<name>.so
#   on the box:  bash scripts/ab_variants.sh run [footprint] [poses] > gpurun_out/ab.jsonl
# Variants: name:flags, one per line in VARIANTS below.
set -e
cd "$(dirname "$0")/.."
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared"
VARIANTS=(
  "base:"
  "noleafmask:-DHDT_NO_PREFIX_LEAF_MASK=1"
  "lowfirst:-DHDT_ANYHIT_LOW_FIRST=1"
  "shblocks10:-DHDT_MIN_BLOCKS_SHADOWS=10"
  "cta64:-DHDT_BLOCK_W=8 -DHDT_BLOCK_H=8 -DHDT_MIN_BLOCKS=24 -DHDT_MIN_BLOCKS_COLORS=32 -DHDT_MIN_BLOCKS_COLORS_RECORDED=32"
  "cta256:-DHDT_BLOCK_W=16 -DHDT_BLOCK_H=16 -DHDT_MIN_BLOCKS=6 -DHDT_MIN_BLOCKS_COLORS=8 -DHDT_MIN_BLOCKS_COLORS_RECORDED=8"
  "search3:-DHDT_COLOR_SEARCH_K=3"
  "colhoist:-DHDT_COLORS_HOIST=1"
)
mkdir -p ab_libs
if [ "$1" = "build" ]; then
    for v in "${VARIANTS[@]}"; do
        n=${v%%:*}; f=${v#*:}
        nvcc $FLAGS $f -o ab_libs/lib_$n.so hashdag_b200/csrc/hdt_tracer.cu -ldl &
    done
    wait
    ls -la ab_libs
else
    libs=""
    for v in "${VARIANTS[@]}" "base:"; do libs="$libs ab_libs/lib_${v%%:*}.so"; done
    python scripts/ab_bench.py --footprint ${2:-14} --poses ${3:-16} $libs 2>/dev/null | grep "^{"
fi
