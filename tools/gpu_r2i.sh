#!/bin/bash
set -u
OUT=gpurun_out/r2i
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
timeout 600 python tools/accuracy_match.py fmg 480 640 > $OUT/acc_fmg_480.json 2> $OUT/acc.err
echo "fmg480 rc=$?" >> $OUT/log.txt
timeout 900 python tools/accuracy_match.py fmg 1080 1920 > $OUT/acc_fmg_1080.json 2>> $OUT/acc.err
echo "fmg1080 rc=$?" >> $OUT/log.txt
timeout 600 python tools/accuracy_match.py hs 480 640 > $OUT/acc_hs_480.json 2>> $OUT/acc.err
echo "hs480 rc=$?" >> $OUT/log.txt
timeout 300 python tools/accuracy_match.py fmg 120 160 > $OUT/acc_fmg_120.json 2>> $OUT/acc.err
timeout 300 python tools/accuracy_match.py hs 120 160 > $OUT/acc_hs_120.json 2>> $OUT/acc.err
echo done >> $OUT/log.txt
