#!/bin/bash
set -u
OUT=gpurun_out/r2p
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_bands.py -q -x > $OUT/pytest_bands.txt 2>&1
echo "bands rc=$?" >> $OUT/log.txt
timeout 900 python -m pytest tests/test_gpu_sibling_drivers.py -q > $OUT/pytest_siblings.txt 2>&1
echo "siblings rc=$?" >> $OUT/log.txt
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_bands.py --deselect tests/test_gpu_sibling_drivers.py > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
