#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu < /dev/null 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" < /dev/null 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 < /dev/null > gpurun_out/bench_k1.json 2> gpurun_out/bench_k1.err; echo "bench rc $?"
PDEGPU_GRAPHS=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_flow_r01b.csv python tools/flow_bench.py 16 1 < /dev/null > gpurun_out/ncu_flow.log 2>&1; echo "flow launch list rc $?"
