#!/bin/bash
# round 2, call A: first contact of the generation-3 line kernel with a GPU
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
# 1. small checks against generation 0 (every family), each bounded
for spec in "llin4 64 96 3" "elin4 64 96 3" "disp 64 96 2" "pde4 64 96 3" "llin8 64 96 2" "llin4 37 53 2" "llin4 480 640 3" "llin4 203 270 2" "elin4 270 360 4" "llin4 120 800 2"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 3 --check --reps 2 --tag check >> $OUT/checks.jsonl 2>> $OUT/checks.err
  echo "check $spec rc=$?" >> $OUT/log.txt
done
# 2. the sweep test file
timeout 900 python -m pytest tests/test_gpu_sweeps.py -x -q > $OUT/pytest_sweeps.txt 2>&1
echo "pytest rc=$?" >> $OUT/log.txt
# 3. geometry sweep on the bench workload
run() { timeout 120 env "$@" python tools/tl_bench.py --fam llin4 --nr 480 --nc 640 --batch 64 --iter 4 --reps 5 >> $OUT/sweep.jsonl 2>> $OUT/sweep.err; }
run PDEGPU_ALR_GEN=2
run PDEGPU_ALR_GEN=3
run PDEGPU_TL_BL=4 PDEGPU_TL_R=16 PDEGPU_TL_D=3
run PDEGPU_TL_BL=4 PDEGPU_TL_R=16 PDEGPU_TL_D=4
run PDEGPU_TL_BL=4 PDEGPU_TL_R=12 PDEGPU_TL_D=2
run PDEGPU_TL_BL=8 PDEGPU_TL_R=24 PDEGPU_TL_D=6
run PDEGPU_TL_BL=8 PDEGPU_TL_R=24 PDEGPU_TL_D=8
run PDEGPU_TL_BL=8 PDEGPU_TL_R=16 PDEGPU_TL_D=4
run PDEGPU_TL_BL=8 PDEGPU_TL_R=32 PDEGPU_TL_D=6
run PDEGPU_TL_BL=8 PDEGPU_TL_R=24 PDEGPU_TL_D=4 PDEGPU_TL_K=4
run PDEGPU_TL_BL=8 PDEGPU_TL_R=24 PDEGPU_TL_D=4 PDEGPU_TL_K=6
run PDEGPU_TL_BL=8 PDEGPU_TL_R=24 PDEGPU_TL_D=4 PDEGPU_TL_NCW=8
run PDEGPU_TL_BL=8 PDEGPU_TL_R=24 PDEGPU_TL_D=4 PDEGPU_TL_NCW=11
run PDEGPU_TL_FUSE_FINAL=0
# other families / shapes at default geometry
for spec in "elin4 480 640 64" "disp 480 640 64" "pde4 480 640 64" "llin8 480 640 32" "llin4 640 480 64" "llin4 270 360 64" "llin4 800 800 32"; do
  set -- $spec
  timeout 120 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --reps 5 >> $OUT/shapes.jsonl 2>> $OUT/shapes.err
  timeout 120 env PDEGPU_ALR_GEN=2 python tools/tl_bench.py --fam $1 --nr $2 --nc $3 --batch $4 --iter 4 --reps 5 --tag gen2 >> $OUT/shapes.jsonl 2>> $OUT/shapes.err
done
echo done >> $OUT/log.txt
