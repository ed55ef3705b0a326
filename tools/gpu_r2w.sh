#!/bin/bash
set -u
OUT=gpurun_out/r2w
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.txt 2>&1
echo "smoke rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
