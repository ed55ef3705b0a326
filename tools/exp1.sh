#!/bin/bash
cd /root/repo
timeout 400 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu < /dev/null 2>&1 | tail -6
timeout 400 python bench.py --steps 5 --warmup 3 < /dev/null > gpurun_out/bench_e1.json 2> gpurun_out/bench_e1.err; echo "bench rc $?"
PDEGPU_GRAPHS=0 timeout 400 python bench.py --steps 5 --warmup 3 < /dev/null > gpurun_out/bench_e1_nograph.json 2> gpurun_out/bench_e1_nograph.err; echo "bench nograph rc $?"
