#!/bin/bash
set -u
OUT=gpurun_out/r3a
mkdir -p $OUT
C1="python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch 148 --iter 1 --reps 1"
PDEGPU_ORDER=reference $C1 > $OUT/plain1.log 2>&1 &&
PDEGPU_ORDER=reference timeout 600 ncu --set full --clock-control none --import-source on -k regex:lex_pass -s 2 -c 2 -o $OUT/prof_lex $C1 > $OUT/ncu1.log 2>&1
echo "ncu lex rc=$?" >> $OUT/log.txt
C2="python tools/tl_bench.py --fam llin4 --nr 48 --nc 64 --batch 16 --iter 4 --reps 1"
$C2 > $OUT/plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tline_small -s 2 -c 1 -o $OUT/prof_small $C2 > $OUT/ncu2.log 2>&1
echo "ncu small rc=$?" >> $OUT/log.txt
C3="python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch 64 --iter 1 --reps 1"
$C3 > $OUT/plain3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tline_prep\|tline_final -s 2 -c 2 -o $OUT/prof_prep $C3 > $OUT/ncu3.log 2>&1
echo "ncu prep rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
