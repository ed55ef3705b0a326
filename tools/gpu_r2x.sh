#!/bin/bash
set -u
OUT=gpurun_out/r2x
mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err
echo "bench4 rc=$?" >> $OUT/log.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --workload band --band-transport nccl --steps 5 --warmup 3 > $OUT/band_4gpu_nccl.json 2> $OUT/band_nccl.err
echo "band nccl rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
