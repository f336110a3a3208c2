#!/bin/bash
# Build library variants from compile-time switches and time each with scripts/ab_bench.py (per-pass kernel ms, oracle check).
#   here:        bash scripts/ab_variants.sh build            # nvcc only, writes ab_libs/lib_<name>.so
#   on the box:  gpurun -- 'bash scripts/ab_variants.sh run'  # one JSON line per variant
# Variants ready to measure (default off, SASS of the default build unchanged):
#   leaflate   -DHDT_LEAF_LATE=1      the 64-bit leaf's first reader placed behind the mask arithmetic
#   colhoist   -DHDT_COLORS_HOIST=1   colour walk: next node's pointer load issued beside the colour-tree load
set -e
cd "$(dirname "$0")/.."
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared"
VARIANTS=("base:" "leaflate:-DHDT_LEAF_LATE=1" "colhoist:-DHDT_COLORS_HOIST=1" "both:-DHDT_LEAF_LATE=1 -DHDT_COLORS_HOIST=1")
mkdir -p ab_libs
if [ "$1" = "build" ]; then
    for v in "${VARIANTS[@]}"; do
        n=${v%%:*}; f=${v#*:}
        nvcc $FLAGS $f -o ab_libs/lib_$n.so hashdag_b200/csrc/hdt_tracer.cu
        echo "built ab_libs/lib_$n.so ($f)"
    done
else
    for v in "${VARIANTS[@]}" "base:"; do
        n=${v%%:*}
        HDT_LIB=$PWD/ab_libs/lib_$n.so AB_RESOLVED=1 timeout 300 python scripts/ab_bench.py ${2:-14} ${3:-16} 2>/dev/null | grep "^{"
    done
fi
