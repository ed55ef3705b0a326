#!/bin/bash
set -u
OUT=gpurun_out/r2l
mkdir -p $OUT
for spec in "llin4 15 20 16" "llin4 27 36 16" "llin4 48 64 16" "elin4 36 48 3" "disp 48 64 5" "pde4 33 41 4" "llin4 64 86 16" "pde8 37 53 3" "pde8 480 640 2"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --omega 1.5 --check --reps 5 --tag small >> $OUT/checks.jsonl 2>> $OUT/checks.err
  echo "check $spec rc=$?" >> $OUT/log.txt
  PDEGPU_TL_SMALL=0 timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --omega 1.5 --reps 5 --tag nosmall >> $OUT/checks.jsonl 2>> $OUT/checks.err
done
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
PDEGPU_GRAPHS=0 PDEGPU_TL_SMALL=0 timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_profile_nosmall.txt 2>&1
PDEGPU_GRAPHS=0 timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_profile.txt 2>&1
timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_graphs.txt 2>&1
echo "flow rc=$?" >> $OUT/log.txt
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt
timeout 300 python bench.py --steps 2 --warmup 1 --flow-batch 0 --fmg-pairs 0 --sweep-legs 0 > $OUT/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tline_pass -s 8 -c 2 -o $OUT/prof_tline python bench.py --steps 2 --warmup 1 --flow-batch 0 --fmg-pairs 0 --sweep-legs 0 > $OUT/ncu.log 2>&1
echo "ncu rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
