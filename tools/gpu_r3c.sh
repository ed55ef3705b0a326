#!/bin/bash
set -u
OUT=gpurun_out/r3c
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_reference_order.py tests/test_gpu_fullsize.py -q > $OUT/pytest_lex.txt 2>&1
echo "lex rc=$?" >> $OUT/log.txt
for spec in "llin4 480 640 64 1" "llin4 480 640 148 1" "elin4 1080 1920 8 2" "llin4 480 640 64 2"; do
  set -- $spec
  PDEGPU_ORDER=reference timeout 200 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --omega 1.0 --solver $5 --reps 2 --tag lex$5 >> $OUT/lex_bench.jsonl 2>> $OUT/lex_bench.err
done
echo done >> $OUT/log.txt
