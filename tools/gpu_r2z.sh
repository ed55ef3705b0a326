#!/bin/bash
OUT=gpurun_out/r2z
mkdir -p $OUT
timeout 900 python tools/disp_diag.py > $OUT/disp_diag.txt 2>&1
echo "rc=$?" >> $OUT/log.txt
