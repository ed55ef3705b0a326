#!/bin/bash
set -u
OUT=gpurun_out/r2u
mkdir -p $OUT
for spec in "llin4 480 640 64" "llin4 480 640 148" "elin4 480 640 64" "disp 256 300 5" "pde4 512 128 3" "llin4 1080 1920 8" "llin4 16384 2048 1"; do
  set -- $spec
  timeout 200 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --omega 1.0 --solver 1 --check --reps 5 --tag window >> $OUT/point.jsonl 2>> $OUT/point.err
  echo "check $spec rc=$?" >> $OUT/log.txt
done
timeout 900 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_bands.py -q -x > $OUT/pytest_point.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
