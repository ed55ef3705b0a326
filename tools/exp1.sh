#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu < /dev/null 2>&1 | tail -3
timeout 500 python bench.py < /dev/null > gpurun_out/bench_i1.json 2> gpurun_out/bench_i1.err; echo "bench rc $?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 < /dev/null > gpurun_out/bench_i1_ref.json 2> gpurun_out/bench_i1_ref.err; echo "ref rc $?"
timeout 300 python bench.py --solver 1 --flow-batch 0 --fmg-pairs 0 < /dev/null > gpurun_out/bench_i1_point.json 2> gpurun_out/bench_i1_point.err; echo "point rc $?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" < /dev/null 2>&1 | tail -2
