#!/bin/bash
cd /root/repo
timeout 400 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_pipeline.py -x -q -m gpu < /dev/null 2>&1 | tail -3
P="timeout 120 python tools/w2_probe.py"
for sz in "64 270 360" "16 270 360" "16 203 270" "16 66 88"; do echo "== batch nr nc = $sz"; $P $sz < /dev/null 2>&1 | grep "alr_\|transpose"; done
timeout 300 python bench.py --steps 5 --warmup 3 < /dev/null > gpurun_out/bench_h2.json 2> gpurun_out/bench_h2.err; echo "bench rc $?"
