#!/bin/bash
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
for spec in "llin4 120 800 2" "llin4 480 640 1" "llin4 800 800 1" "llin8 96 64 2" "pde4 37 53 3" "disp 203 270 2" "elin4 640 480 2"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 3 --check --reps 2 --tag check >> $OUT/checks.jsonl 2>> $OUT/checks.err
  echo "check $spec rc=$?" >> $OUT/log.txt
done
timeout 900 python -m pytest tests/test_gpu_sweeps.py -x -q > $OUT/pytest_sweeps.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
P=pde-based-image-processing_b200/libpdegpu_probe.so
PDEGPU_LIB=$P PDEGPU_TL_FUSE_FINAL=0 timeout 120 python tools/tl_probe.py >> $OUT/probe.txt 2>&1
PDEGPU_LIB=$P PDEGPU_TL_FUSE_FINAL=0 PDEGPU_TL_K=4 timeout 120 python tools/tl_probe.py >> $OUT/probe.txt 2>&1
PDEGPU_LIB=$P PDEGPU_TL_FUSE_FINAL=0 PDEGPU_TL_NCW=6 timeout 120 python tools/tl_probe.py >> $OUT/probe.txt 2>&1
PDEGPU_LIB=$P PDEGPU_TL_FUSE_FINAL=0 PDEGPU_TL_BL=4 PDEGPU_TL_R=16 PDEGPU_TL_D=3 timeout 120 python tools/tl_probe.py >> $OUT/probe.txt 2>&1
PDEGPU_TL_FUSE_FINAL=0 timeout 200 python tools/tl_bench.py --iter 1 --reps 1 > $OUT/plain.json 2> $OUT/plain.err && \
PDEGPU_TL_FUSE_FINAL=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tline_pass -s 4 -c 2 -o $OUT/tline_pass python tools/tl_bench.py --iter 1 --reps 1 > $OUT/ncu.log 2>&1
echo "ncu rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
