#!/bin/bash
set -u
OUT=gpurun_out/r2q
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_bands.py tests/test_gpu_sibling_drivers.py -q > $OUT/pytest_bands.txt 2>&1
echo "bands+siblings rc=$?" >> $OUT/log.txt
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
