#!/bin/bash
set -u
OUT=gpurun_out/r2t
mkdir -p $OUT
for spec in "llin4 480 640 64" "elin4 480 640 8" "disp 256 300 5" "pde4 512 128 3" "llin4 1080 1920 4" "llin4 64 80 7"; do
  set -- $spec
  timeout 200 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --omega 1.0 --solver 1 --check --reps 5 --tag window >> $OUT/point.jsonl 2>> $OUT/point.err
  echo "check $spec rc=$?" >> $OUT/log.txt
done
PDEGPU_POINT_WINDOW=0 timeout 200 python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch 64 --iter 4 --omega 1.0 --solver 1 --reps 5 --tag tiles >> $OUT/point.jsonl 2>> $OUT/point.err
timeout 200 python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch 148 --iter 4 --omega 1.0 --solver 1 --reps 5 --tag window148 >> $OUT/point.jsonl 2>> $OUT/point.err
timeout 900 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_bands.py tests/test_gpu_parity_fullsize.py -q -x > $OUT/pytest_point.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
