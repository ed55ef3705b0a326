#!/bin/bash
# what the driver does at round end: GPU tests, smoke, both bench arms
set -u
OUT=gpurun_out/r4e
mkdir -p $OUT
date +%s > $OUT/t0
timeout 600 python -m pytest tests -x -q -m gpu < /dev/null > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt; date +%s > $OUT/t1
timeout 600 python bench.py --steps 20 --warmup 3 < /dev/null > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/log.txt; date +%s > $OUT/t2
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 < /dev/null > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "ref rc=$?" >> $OUT/log.txt; date +%s > $OUT/t3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" < /dev/null > $OUT/smoke.txt 2>&1
echo "smoke rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
PDEGPU_GRAPHS=0 timeout 90 python tools/flow_bench.py 64 2 > $OUT/flow64_profile.txt 2>&1
