#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_sweeps.py -x -q -m gpu < /dev/null 2>&1 | tail -3
P="timeout 120 python tools/w2_probe.py 64"
D=pde-based-image-processing_b200
echo "== point async (3 CTAs/SM)"; W2_SOLVER=1 $P < /dev/null 2>&1 | grep rb_tile
echo "== point async (2 CTAs/SM, no spills)"; W2_SOLVER=1 PDEGPU_LIB=$D/libpdegpu_pta2.so $P < /dev/null 2>&1 | grep rb_tile
echo "== point register-staged"; W2_SOLVER=1 PDEGPU_POINT_ASYNC=0 $P < /dev/null 2>&1 | grep rb_tile
