#!/bin/bash
# On the GPU box: compute-sanitizer over parity / option tests (memcheck) and over the cases that exercise the
# shared-memory machinery (racecheck: counting-sort histograms, phase barriers, cp.async record pipeline).
TAG=${1:-san}; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_options.py tests/test_gpu_parity.py -m gpu -x -q \
   -k "not full_size and not config and not exhaustive and not device_math" > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_options.py -m gpu -x -q -k "max_depth or more_than_32" > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $OUT/${TAG}_racecheck.log
