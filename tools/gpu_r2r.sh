#!/bin/bash
set -u
OUT=gpurun_out/r2r
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_bands.py -q > $OUT/pytest_bands.txt 2>&1
echo "bands rc=$?" >> $OUT/log.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
echo "bench2 rc=$?" >> $OUT/log.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload band --band-transport nccl --steps 5 --warmup 3 > $OUT/band_2gpu_nccl.json 2> $OUT/band_nccl.err
echo "band nccl rc=$?" >> $OUT/log.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload band --band-transport p2p --steps 5 --warmup 3 > $OUT/band_2gpu_p2p.json 2> $OUT/band_p2p.err
echo "band p2p rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
