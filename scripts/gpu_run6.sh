#!/bin/bash
OUT=gpurun_out/run7
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
nvidia-smi --query-gpu=index,name --format=csv
echo "== cpp shim test"; timeout 600 python -m pytest tests/test_gpu_cpp_shim.py -m gpu -q 2>&1 | tail -5
echo "== N=2 bench"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 8 --footprint-log2 13 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; cat $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
echo "== N=1 bench 4K"; timeout 900 python bench.py --gpus 1 --steps 32 --warmup 8 --footprint-log2 13 --width 3840 --height 2160 --no-cpu-baseline > $OUT/bench_n1_4k.json 2> $OUT/bench_n1_4k.err; cat $OUT/bench_n1_4k.json; tail -3 $OUT/bench_n1_4k.err
echo "== N=2 ref arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 --footprint-log2 13 2>&1 | tail -2
