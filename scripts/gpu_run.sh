#!/bin/bash
# Everything that is run on the GPU box goes through this one script:  gpurun -- 'bash scripts/gpu_run.sh <stage> [args]'
# Results land in gpurun_out/ (merged back by gpurun); stages are independent.
#   tests                     pytest -m gpu
#   ab [footprint] [poses]    A/B of the library variants in ab_libs/ (scripts/ab_variants.sh build, here)
#   bench [bench.py args]     one bench line (+ clocks)
#   launches                  ncu launch list of a short bench run (per-launch device times, cold caches)
#   ncu                       ncu --set full of the three per-pixel kernels of a short bench run
#   smoke                     __graft_entry__.smoke()
#   colorleaf                 colour-leaf rebuild: its tests, scripts/bench_color_leaf.py, ncu --set full of its kernels
#   sanitize <tool> <pytest args>   compute-sanitizer --tool memcheck|racecheck over a pytest selection
#   golden <generator.py>     a tests/golden/make_*.py fixture generator (reference harness) -> gpurun_out/golden/
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${HDT_RUN_TAG:-r2}
stage=$1; shift
case "$stage" in
tests)
    python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.log ;;
ab)
    bash scripts/ab_variants.sh run "$@" | tee gpurun_out/${TAG}_ab.jsonl ;;
bench)
    python bench.py "$@" 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json ;;
launches)
    ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
        python bench.py --steps 8 --warmup 3 --poses 8 --frames-in-flight 1 --no-cpu-baseline --no-ref-cuda "$@" > gpurun_out/${TAG}_launches.log 2>&1 ;;
ncu)
    ncu --set full --metrics l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum --clock-control none --import-source on -k regex:'trace_(paths|shadows|colors)' -s 22 -c 6 -f -o gpurun_out/${TAG}_prof \
        python bench.py --steps 4 --warmup 2 --poses 8 --frames-in-flight 1 --no-cpu-baseline --no-ref-cuda "$@" > gpurun_out/${TAG}_ncu.log 2>&1
    ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null ;;
smoke)
    python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log ;;
colorleaf)
    timeout 900 python -m pytest tests/test_gpu_color_leaf.py -x -q 2>&1 | tail -30 | tee gpurun_out/${TAG}_cl_pytest.log
    timeout 600 python scripts/bench_color_leaf.py > gpurun_out/${TAG}_color_leaf_bench.json 2> gpurun_out/${TAG}_color_leaf_bench.err
    tail -c 1500 gpurun_out/${TAG}_color_leaf_bench.json
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:color_ -c 4 -f -o gpurun_out/${TAG}_cl_prof python scripts/bench_color_leaf.py --reps 1 > gpurun_out/${TAG}_cl_ncu.log 2>&1
    ncu -i gpurun_out/${TAG}_cl_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_cl_prof_raw.csv 2>/dev/null ;;
sanitize)
    tool=$1; shift
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest "$@" -x -q 2>&1 | tail -40 | tee gpurun_out/${TAG}_sanitizer_${tool}.log ;;
golden)
    timeout 900 python "$1" gpurun_out/golden 2>&1 | tail -20 | tee gpurun_out/${TAG}_golden.log ;;
*)
    echo "unknown stage $stage"; exit 2 ;;
esac
