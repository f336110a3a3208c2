#!/bin/bash
OUT=gpurun_out/n2
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python bench.py --gpus 1 --steps 32 --warmup 8 --width 3840 --height 2160 --no-cpu-baseline > $OUT/n1_4k.json 2> $OUT/n1.err; cat $OUT/n1_4k.json
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 32 --warmup 8 > $OUT/n$N.json 2> $OUT/n$N.err; cat $OUT/n$N.json; grep -i "error\|Traceback" $OUT/n$N.err | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+N)) bench.py --impl reference --gpus $N --steps 4 --warmup 1 > $OUT/ref_n$N.json 2> $OUT/ref_n$N.err; cat $OUT/ref_n$N.json | cut -c1-400
