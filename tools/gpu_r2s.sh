#!/bin/bash
set -u
OUT=gpurun_out/r2s
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_driver_steps.py tests/test_gpu_pipeline.py tests/test_gpu_reference_order.py -q -x > $OUT/pytest_fused.txt 2>&1
echo "fused rc=$?" >> $OUT/log.txt
PDEGPU_GRAPHS=0 timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_profile.txt 2>&1
timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_graphs.txt 2>&1
timeout 300 python tools/flow_bench.py 64 3 > $OUT/flow_graphs64.txt 2>&1
PDEGPU_FUSE=0 timeout 300 python tools/flow_bench.py 16 3 > $OUT/flow_graphs_nofuse.txt 2>&1
echo "flow rc=$?" >> $OUT/log.txt
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
echo done >> $OUT/log.txt
