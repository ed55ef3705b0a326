#!/bin/bash
set -u
OUT=gpurun_out/r2k
mkdir -p $OUT
for spec in "llin4 480 640 3" "llin8 96 64 2" "llin8 203 270 2" "pde8 37 53 3" "pde8 131 67 2" "pde8 480 640 2" "pde4 131 67 3"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 3 --check --reps 2 --tag check >> $OUT/checks.jsonl 2>> $OUT/checks.err
  echo "check $spec rc=$?" >> $OUT/log.txt
done
for spec in "pde8 1080 1920 1" "llin8 1080 900 1" "pde8 2160 4096 1"; do
  set -- $spec
  timeout 300 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --omega 1.0 --check --reps 2 --tag long >> $OUT/long.jsonl 2>> $OUT/long.err
  echo "long $spec rc=$?" >> $OUT/log.txt
  PDEGPU_ALR_GEN=2 timeout 300 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --omega 1.0 --reps 2 --tag long-gen2 >> $OUT/long.jsonl 2>> $OUT/long.err
done
timeout 1200 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_parity_fullsize.py -x -q > $OUT/pytest_sweeps.txt 2>&1
echo "pytest sweeps rc=$?" >> $OUT/log.txt
timeout 900 python -m pytest tests/test_gpu_configs_fixtures.py -x -q -s > $OUT/pytest_fixtures.txt 2>&1
echo "fixtures rc=$?" >> $OUT/log.txt
timeout 300 python tools/accuracy_match.py hs 480 640 > $OUT/acc_hs_480.json 2>> $OUT/acc.err
timeout 300 python tools/accuracy_match.py hs 120 160 > $OUT/acc_hs_120.json 2>> $OUT/acc.err
PDEGPU_GRAPHS=0 timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_profile.txt 2>&1
timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_graphs.txt 2>&1
timeout 200 python tools/tl_bench.py --fam pde8 --nr 2160 --nc 4096 --batch 4 --iter 1 --reps 5 >> $OUT/shapes.jsonl 2>> $OUT/shapes.err
PDEGPU_ALR_GEN=2 timeout 200 python tools/tl_bench.py --fam pde8 --nr 2160 --nc 4096 --batch 4 --iter 1 --reps 5 --tag gen2 >> $OUT/shapes.jsonl 2>> $OUT/shapes.err
echo done >> $OUT/log.txt
