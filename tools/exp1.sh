#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_sweeps.py -x -q -m gpu < /dev/null 2>&1 | tail -3
P="timeout 120 python tools/w2_probe.py 64"
D=pde-based-image-processing_b200
for g in 2 1 4; do echo "== G=$g"; PDEGPU_W2_G=$g $P < /dev/null 2>&1 | grep alr_window; done
echo "== G=2 probe"; PDEGPU_LIB=$D/libpdegpu_probe.so $P < /dev/null 2>&1 | tail -22 | head -14
echo "== G=4 probe"; PDEGPU_W2_G=4 PDEGPU_LIB=$D/libpdegpu_probe.so $P < /dev/null 2>&1 | tail -22 | head -14
echo "== G=2 NA=10 NS=2... (NA+NS<=12)"; PDEGPU_W2_NA=10 PDEGPU_W2_NS=2 $P < /dev/null 2>&1 | grep alr_window
echo "== G=2 NA=8 NS=4 R=24 D=4 NBUF10"; PDEGPU_W2_R=24 PDEGPU_W2_D=4 PDEGPU_W2_NBUF=10 $P < /dev/null 2>&1 | grep alr_window
echo "== G=4 R=24 D=4 NBUF10"; PDEGPU_W2_G=4 PDEGPU_W2_R=24 PDEGPU_W2_D=4 PDEGPU_W2_NBUF=10 $P < /dev/null 2>&1 | grep alr_window
echo "== G=3 NA=9 NS=3"; PDEGPU_W2_G=3 PDEGPU_W2_NA=9 PDEGPU_W2_NS=3 $P < /dev/null 2>&1 | grep alr_window
