#!/bin/bash
set -u
OUT=gpurun_out/r4b
mkdir -p $OUT
PDEGPU_LANES=64 timeout 200 python tools/fmg_lanes.py 64 2 >> $OUT/fmg.txt 2>> $OUT/fmg.err
PDEGPU_LANES=128 timeout 200 python tools/fmg_lanes.py 128 2 >> $OUT/fmg.txt 2>> $OUT/fmg.err
CUDA_DEVICE_MAX_CONNECTIONS=32 PDEGPU_LANES=128 timeout 200 python tools/fmg_lanes.py 128 2 >> $OUT/fmg.txt 2>> $OUT/fmg.err
PDEGPU_LANES=128 timeout 200 python tools/fmg_lanes.py 128 2 fast >> $OUT/fmg.txt 2>> $OUT/fmg.err
timeout 100 python tools/flow_bench.py 16 3 > $OUT/flow16.txt 2>&1
echo done >> $OUT/log.txt
