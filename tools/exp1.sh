#!/bin/bash
cd /root/repo
timeout 500 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu < /dev/null 2>&1 | tail -12
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_sweeps.py -x -q -m gpu -k "not 1080 and not 480" < /dev/null > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?"; tail -5 gpurun_out/sanitizer_memcheck.log
