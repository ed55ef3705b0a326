#!/bin/bash
# what the driver does at round end, in one gpurun call: GPU tests, smoke, one bench line
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu < /dev/null 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" < /dev/null 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 < /dev/null > gpurun_out/bench_k1.json 2> gpurun_out/bench_k1.err; echo "bench rc $?"
