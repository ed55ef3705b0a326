#!/bin/bash
set -u
OUT=gpurun_out/r2o
mkdir -p $OUT
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "bench ref rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
