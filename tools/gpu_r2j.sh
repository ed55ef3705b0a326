#!/bin/bash
set -u
OUT=gpurun_out/r2j
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_configs_fixtures.py -x -q -s > $OUT/pytest_fixtures.txt 2>&1
echo "fixtures rc=$?" >> $OUT/log.txt
timeout 300 python tools/accuracy_match.py hs 480 640 > $OUT/acc_hs_480.json 2>> $OUT/acc.err
timeout 300 python tools/accuracy_match.py hs 120 160 > $OUT/acc_hs_120.json 2>> $OUT/acc.err
PDEGPU_GRAPHS=0 timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_profile.txt 2>&1
timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_graphs.txt 2>&1
echo done >> $OUT/log.txt
