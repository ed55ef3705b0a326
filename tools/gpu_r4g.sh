#!/bin/bash
set -u
OUT=gpurun_out/r4g
mkdir -p $OUT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --batch512 0 --flow-ref-batch 0 --fmg-pairs 0 --sweep-legs 0 --flow-batch 16 --band-leg 0 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
echo "bench2 rc=$?" >> $OUT/log.txt
