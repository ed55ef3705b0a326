#!/bin/bash
set -u
OUT=gpurun_out/r2d
mkdir -p $OUT
for spec in "llin4 64 96 3" "elin4 64 96 3" "disp 64 96 2" "pde4 37 53 3" "llin8 96 64 2" "llin4 37 53 2" "llin4 480 640 1" "llin4 480 640 3" "llin4 203 270 2" "elin4 270 360 4" "llin4 120 800 2" "llin4 800 800 1" "elin4 640 480 2" "llin4 540 100 2" "llin8 203 270 2"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 3 --check --reps 2 --tag check >> $OUT/checks.jsonl 2>> $OUT/checks.err
  echo "check $spec rc=$?" >> $OUT/log.txt
done
timeout 900 python -m pytest tests/test_gpu_sweeps.py -x -q > $OUT/pytest_sweeps.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
run() { timeout 120 env "$@" python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch 64 --iter 4 --reps 5 >> $OUT/sweep.jsonl 2>> $OUT/sweep.err; }
run PDEGPU_ALR_GEN=3
run PDEGPU_TL_NCW=8
run PDEGPU_TL_NCW=6
run PDEGPU_TL_BL=4 PDEGPU_TL_R=16 PDEGPU_TL_D=3
run PDEGPU_TL_BL=8 PDEGPU_TL_R=24 PDEGPU_TL_D=6
run PDEGPU_TL_BL=8 PDEGPU_TL_R=32 PDEGPU_TL_D=6
run PDEGPU_TL_BL=8 PDEGPU_TL_R=32 PDEGPU_TL_D=8
run PDEGPU_TL_BL=8 PDEGPU_TL_R=40 PDEGPU_TL_D=8
run PDEGPU_TL_BL=8 PDEGPU_TL_R=24 PDEGPU_TL_D=4 PDEGPU_TL_K=4
run PDEGPU_TL_BL=8 PDEGPU_TL_R=24 PDEGPU_TL_D=4 PDEGPU_TL_K=6
run PDEGPU_TL_BL=8 PDEGPU_TL_R=16 PDEGPU_TL_D=4
P=pde-based-image-processing_b200/libpdegpu_probe.so
PDEGPU_LIB=$P timeout 120 python tools/tl_probe.py >> $OUT/probe.txt 2>&1
PDEGPU_LIB=$P PDEGPU_TL_NCW=8 timeout 120 python tools/tl_probe.py >> $OUT/probe.txt 2>&1
PDEGPU_LIB=$P PDEGPU_TL_BL=8 PDEGPU_TL_R=32 PDEGPU_TL_D=6 timeout 120 python tools/tl_probe.py >> $OUT/probe.txt 2>&1
for spec in "elin4 480 640 64" "disp 480 640 64" "pde4 480 640 64" "llin8 480 640 32" "llin4 270 360 64" "llin4 800 800 32" "llin4 120 160 64" "llin4 60 80 64" "llin4 544 544 64"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --reps 5 >> $OUT/shapes.jsonl 2>> $OUT/shapes.err
done
timeout 200 python tools/tl_bench.py --iter 1 --reps 1 > $OUT/plain.json 2> $OUT/plain.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tline_ -s 5 -c 4 -o $OUT/tline python tools/tl_bench.py --iter 1 --reps 1 > $OUT/ncu.log 2>&1
echo "ncu rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
