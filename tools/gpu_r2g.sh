#!/bin/bash
set -u
OUT=gpurun_out/r2g
mkdir -p $OUT
# regression on short lines
for spec in "llin4 480 640 3" "llin4 203 270 2" "elin4 64 96 3" "disp 203 270 3" "llin8 96 64 2" "pde4 131 67 3"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 3 --check --reps 2 --tag check >> $OUT/checks.jsonl 2>> $OUT/checks.err
  echo "check $spec rc=$?" >> $OUT/log.txt
done
# long lines: difference to generation 0 (exact zebra: cuts make a small difference), residual histories of both generations
for spec in "llin4 1080 1920 1" "elin4 1080 1920 2" "llin4 810 1000 2" "disp 2160 4096 1" "pde4 1080 900 2" "llin4 1920 1080 1"; do
  set -- $spec
  timeout 300 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --omega 1.0 --check --resid --reps 2 --tag long >> $OUT/long.jsonl 2>> $OUT/long.err
  echo "long $spec rc=$?" >> $OUT/log.txt
  PDEGPU_ALR_GEN=2 timeout 300 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --omega 1.0 --resid --reps 2 --tag long-gen2 >> $OUT/long.jsonl 2>> $OUT/long.err
done
timeout 600 python -m pytest tests/test_gpu_sweeps.py -x -q > $OUT/pytest_sweeps.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
# throughput at the large shapes
for spec in "elin4 1080 1920 8" "llin4 1080 1920 8" "disp 2160 4096 4" "pde4 2160 4096 4" "elin4 2160 4096 2"; do
  set -- $spec
  timeout 200 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --reps 3 >> $OUT/shapes.jsonl 2>> $OUT/shapes.err
  PDEGPU_ALR_GEN=2 timeout 200 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --reps 3 --tag gen2 >> $OUT/shapes.jsonl 2>> $OUT/shapes.err
done
run() { timeout 120 env "$@" python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch 64 --iter 4 --reps 5 >> $OUT/sweep.jsonl 2>> $OUT/sweep.err; }
run PDEGPU_ALR_GEN=3
run PDEGPU_TL_K=4
run PDEGPU_TL_K=3
run PDEGPU_TL_K=6
run PDEGPU_TL_NCW=5
run PDEGPU_TL_NCW=4 PDEGPU_TL_K=4
echo done >> $OUT/log.txt
