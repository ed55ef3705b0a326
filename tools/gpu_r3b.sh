#!/bin/bash
OUT=gpurun_out/r3b
mkdir -p $OUT
timeout 300 python tools/sanitize.py > $OUT/plain.log 2>&1 &&
PDEGPU_GRAPHS=0 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/log.txt
