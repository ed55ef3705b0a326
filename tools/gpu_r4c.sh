#!/bin/bash
set -u
OUT=gpurun_out/r4c
mkdir -p $OUT
timeout 90 python tools/fmg_lanes.py 64 2 >> $OUT/fmg.txt 2>> $OUT/fmg.err
for b in 16 32 64; do
  timeout 60 python tools/flow_bench.py $b 3 2>&1 | head -2 >> $OUT/flow.txt
done
for b in 64 128; do
  PDEGPU_ORDER=reference timeout 60 python tools/flow_bench.py $b 2 2>&1 | head -2 >> $OUT/flow_ref.txt
done
echo done >> $OUT/log.txt
