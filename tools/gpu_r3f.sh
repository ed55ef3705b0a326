#!/bin/bash
set -u
OUT=gpurun_out/r3f
mkdir -p $OUT
for b in 64 444; do
  PDEGPU_ORDER=reference timeout 200 python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch $b --iter 2 --reps 2 --check --tag lazy >> $OUT/lex.jsonl 2>> $OUT/lex.err
done
PDEGPU_ORDER=reference timeout 200 python tools/tl_bench.py --fam elin4 --nr 1080 --nc 1920 --batch 8 --iter 2 --reps 2 --tag lazy >> $OUT/lex.jsonl 2>> $OUT/lex.err
echo done >> $OUT/log.txt
