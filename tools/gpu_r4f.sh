#!/bin/bash
set -u
OUT=gpurun_out/r4f
mkdir -p $OUT
SKIP="--band-leg 0 --batch512 0 --flow-batch 0 --flow-ref-batch 0 --sweep-legs 0 --fmg-pairs 0"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 $SKIP > $OUT/ncu_bench.log 2>&1
echo "ncu rc=$?" >> $OUT/log.txt
