#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_pipeline.py -x -q -m gpu < /dev/null 2>&1 | tail -15
P="timeout 120 python tools/w2_probe.py 64"
echo "== shipping lib, default";  $P < /dev/null 2>&1 | tail -2
