#!/bin/bash
set -u
OUT=gpurun_out/r2v
mkdir -p $OUT
CMD="python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch 64 --iter 4 --omega 1.0 --solver 1 --reps 2"
$CMD > $OUT/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rb_window -s 2 -c 1 -o $OUT/prof_window $CMD > $OUT/ncu.log 2>&1
echo "ncu rc=$?" >> $OUT/log.txt
