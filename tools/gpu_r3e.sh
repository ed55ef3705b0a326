#!/bin/bash
set -u
OUT=gpurun_out/r3e
mkdir -p $OUT
for cfg in "0 0" "3 4" "3 5" "4 5" "4 6"; do
  set -- $cfg
  for b in 64 148 444; do
    PDEGPU_LEX_KS=$1 PDEGPU_LEX_RL=$2 PDEGPU_ORDER=reference timeout 200 python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch $b --iter 2 --reps 2 --tag "ks$1rl$2" >> $OUT/lex.jsonl 2>> $OUT/lex.err
  done
  PDEGPU_LEX_KS=$1 PDEGPU_LEX_RL=$2 PDEGPU_ORDER=reference timeout 200 python tools/tl_bench.py --fam elin4 --nr 1080 --nc 1920 --batch 8 --iter 2 --reps 2 --tag "ks$1rl$2" >> $OUT/lex.jsonl 2>> $OUT/lex.err
done
echo done >> $OUT/log.txt
