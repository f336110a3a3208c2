#!/bin/bash
# 1/2/4/8-GPU scaling of bench.py on one box (config 3: 3840x2160 split by screen tiles)
OUT=gpurun_out/scale
mkdir -p $OUT
exec > >(tee $OUT/log.txt) 2>&1
F=${F:-14}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python bench.py --gpus 1 --steps 32 --warmup 8 --footprint-log2 $F --no-cpu-baseline > $OUT/n1_1080p.json 2> $OUT/n1.err; cat $OUT/n1_1080p.json
timeout 600 python bench.py --gpus 1 --steps 32 --warmup 8 --footprint-log2 $F --width 3840 --height 2160 --no-cpu-baseline > $OUT/n1_4k.json 2>> $OUT/n1.err; cat $OUT/n1_4k.json
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 32 --warmup 8 --footprint-log2 $F > $OUT/n$N.json 2> $OUT/n$N.err; cat $OUT/n$N.json; grep -i "error\|Traceback" $OUT/n$N.err | head -3
done
