#!/bin/bash
# A/B of colour-leaf rebuild variants (ab_libs/lib_*.so built here with -DHDT_PIECE_* switches): one scripts/bench_color_leaf.py run each.
cd "$(dirname "$0")/.."
for lib in ab_libs/lib_*.so; do
    HDT_LIB=$PWD/$lib timeout 300 python scripts/bench_color_leaf.py --reps 20 --cpu-sample 65536 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); w=d['whole']; print(json.dumps({'lib':'$lib','whole_ms':round(w['kernel_ms'],4),'best':round(w['kernel_ms_best'],4),'edit_ms':round(d['edit']['kernel_ms'],4),'same':d['parity_sample_identical'] and d['edit']['identical_to_oracle']}))"
done
