#!/bin/bash
set -u
OUT=gpurun_out/r3g
mkdir -p $OUT
date +%s > $OUT/t0
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
date +%s > $OUT/t1
timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt
date +%s > $OUT/t2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "ref rc=$?" >> $OUT/log.txt
date +%s > $OUT/t3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.txt 2>&1
echo "smoke rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
