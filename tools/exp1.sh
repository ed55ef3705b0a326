#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu < /dev/null 2>&1 | tail -5
timeout 400 python bench.py < /dev/null > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; echo "bench rc $?"
tail -c 600 gpurun_out/bench_b1.err
