#!/bin/bash
set -u
OUT=gpurun_out/r2h
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity_fullsize.py -x -q --durations=5 > $OUT/pytest_parity.txt 2>&1
echo "parity rc=$?" >> $OUT/log.txt
timeout 2400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity_fullsize.py > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "bench ref rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
