#!/bin/bash
cd /root/repo
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --flow-batch 0 --fmg-pairs 0 < /dev/null > gpurun_out/ncu_b.log 2>&1; echo "launch list rc $?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:alr_window2 -s 8 -c 2 -o gpurun_out/prof_win_g -f python tools/w2_probe.py 64 < /dev/null > gpurun_out/ncu_f.log 2>&1; echo "full rc $?"
W2_SOLVER=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:rb_tile -s 2 -c 1 -o gpurun_out/prof_pt_c -f python tools/w2_probe.py 64 < /dev/null > gpurun_out/ncu_p.log 2>&1; echo "full pt rc $?"
PDEGPU_LIB=pde-based-image-processing_b200/libpdegpu_probe.so timeout 120 python tools/w2_probe.py 64 < /dev/null 2>&1 | tail -22 > gpurun_out/probes_g2.txt
