#!/bin/bash
set -u
OUT=gpurun_out/r4a
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_host_batch.py -x -q -m gpu > $OUT/pytest_batch.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
SKIP="--band-leg 0 --batch512 0 --flow-batch 0 --flow-ref-batch 0 --sweep-legs 0 --fmg-pairs 0"
timeout 300 python bench.py --steps 10 --warmup 3 $SKIP > $OUT/bench_e2e.json 2> $OUT/bench_e2e.err
echo "bench rc=$?" >> $OUT/log.txt
for cfg in "1 1" "2 2" "4 3" "8 3" "4 4" "16 3"; do
  set -- $cfg
  PDEGPU_HOST_CHUNK=$1 PDEGPU_HOST_LANES=$2 timeout 200 python bench.py --steps 5 --warmup 3 $SKIP 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('chunk $1 lanes $2', round(e['value']), round(e['ms_per_step'],2), round(e['pcie_gbs'],1), round(e['single_call']['value']))" >> $OUT/sweep.txt 2>&1
done
echo done >> $OUT/log.txt
