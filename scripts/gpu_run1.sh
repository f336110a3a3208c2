#!/bin/bash
# First GPU trip: parity (oracle / reference / product), golden fixtures, first bench + ncu.
OUT=gpurun_out/run1
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40
echo "== golden"; timeout 300 python tests/golden/make_golden.py $OUT/golden 2>&1 | tail -20
echo "== bench ours F=13"; timeout 600 python bench.py --steps 32 --warmup 8 --footprint-log2 13 > $OUT/bench_ours_f13.json 2> $OUT/bench_ours_f13.err; tail -c 3000 $OUT/bench_ours_f13.json; tail -5 $OUT/bench_ours_f13.err
echo "== bench ref-cuda F=13"; timeout 600 python bench.py --impl reference-cuda --steps 32 --warmup 8 --footprint-log2 13 > $OUT/bench_refcuda_f13.json 2> $OUT/bench_refcuda_f13.err; tail -c 2000 $OUT/bench_refcuda_f13.json; tail -5 $OUT/bench_refcuda_f13.err
echo "== build F=14 timing"; HDS_VERBOSE=1 timeout 600 python -c "
import time,sys
sys.path.insert(0,'.')
from hashdag_b200 import workloads
t=time.time(); s,p=workloads.build_workload(17,14,64); print('F14 build', time.time()-t, s.n_voxels, s.basic.size, s.hash_pool.nbytes/2**20)
"
echo "== ncu launch list (ours)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_ours.csv python bench.py --steps 4 --warmup 2 --footprint-log2 13 --no-cpu-baseline --poses 8 > /dev/null 2>&1; tail -5 $OUT/launches_ours.csv
echo "== ncu full (ours paths+shadows+colors)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_ -s 30 -c 3 -o $OUT/prof_ours python bench.py --steps 4 --warmup 2 --footprint-log2 13 --no-cpu-baseline --poses 8 > /dev/null 2>&1; ls -la $OUT
echo "== ncu full (reference)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_ -s 30 -c 3 -o $OUT/prof_ref python bench.py --impl reference-cuda --steps 4 --warmup 2 --footprint-log2 13 --poses 8 > /dev/null 2>&1; ls -la $OUT
echo done
