#!/bin/bash
OUT=gpurun_out/run8
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
echo "== N=2 bench (pipelined)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 8 --footprint-log2 13 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; cat $OUT/bench_n2.json | cut -c1-400; grep -i "error\|Traceback" -A5 $OUT/bench_n2.err | head -20
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 8 --warmup 3 --footprint-log2 13 --no-cpu-baseline --poses 8 > /dev/null 2>&1; tail -4 $OUT/launches.csv | cut -c1-300
echo "== ncu full v3"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_ -s 33 -c 3 -o $OUT/prof_v3 python bench.py --steps 4 --warmup 2 --footprint-log2 13 --no-cpu-baseline --poses 8 > /dev/null 2>&1; ls -la $OUT | grep ncu
