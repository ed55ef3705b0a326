#!/bin/bash
set -u
OUT=gpurun_out/r3h
mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err
echo "bench8 rc=$?" >> $OUT/log.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --workload band --band-transport nccl --steps 5 --warmup 3 > $OUT/band_8gpu_nccl.json 2> $OUT/band_nccl.err
echo "band nccl rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
