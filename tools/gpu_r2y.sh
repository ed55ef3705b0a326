#!/bin/bash
set -u
OUT=gpurun_out/r2y
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_reference_order.py tests/test_gpu_fullsize.py tests/test_gpu_sibling_drivers.py tests/test_gpu_pipeline.py -q -x > $OUT/pytest_lex.txt 2>&1
echo "lex rc=$?" >> $OUT/log.txt
for spec in "llin4 480 640 64" "llin4 480 640 444" "elin4 1080 1920 8"; do
  set -- $spec
  PDEGPU_ORDER=reference timeout 200 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --reps 3 --tag lex >> $OUT/lex_bench.jsonl 2>> $OUT/lex_bench.err
done
timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
