#!/bin/bash
set -u
OUT=gpurun_out/r4h
mkdir -p $OUT
s=$(date +%s)
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 110 python tools/fmg_lanes.py 64 1 >> $OUT/fmg.txt 2>> $OUT/fmg.err
echo "rc=$? wall=$(( $(date +%s) - s )) s" >> $OUT/fmg.txt
