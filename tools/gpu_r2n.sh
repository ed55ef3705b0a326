#!/bin/bash
set -u
OUT=gpurun_out/r2n
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_reference_order.py tests/test_gpu_configs_fixtures.py -q -s > $OUT/pytest_lex.txt 2>&1
echo "lex rc=$?" >> $OUT/log.txt
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_reference_order.py --deselect tests/test_gpu_configs_fixtures.py > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
for spec in "llin4 480 640 64" "llin4 480 640 444" "elin4 1080 1920 8" "llin4 120 160 148" "disp 2160 4096 2"; do
  set -- $spec
  PDEGPU_ORDER=reference timeout 200 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --reps 3 --tag lex >> $OUT/lex_bench.jsonl 2>> $OUT/lex_bench.err
done
echo done >> $OUT/log.txt
