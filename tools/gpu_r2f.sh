#!/bin/bash
set -u
OUT=gpurun_out/r2f
mkdir -p $OUT
for spec in "llin4 480 640 1" "llin4 480 640 3" "llin4 203 270 2" "llin8 203 270 2" "llin8 480 640 2" "disp 203 270 3" "pde4 131 67 3" "llin4 800 800 1" "llin4 9 8 2" "llin4 8 37 1"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 3 --check --reps 2 --tag check >> $OUT/checks.jsonl 2>> $OUT/checks.err
  echo "check $spec rc=$?" >> $OUT/log.txt
done
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt
for spec in "llin4 800 800 32" "llin4 640 480 64" "llin4 480 640 64"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --reps 5 >> $OUT/shapes.jsonl 2>> $OUT/shapes.err
done
echo done >> $OUT/log.txt
