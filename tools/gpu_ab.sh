#!/bin/bash
# On the GPU box: A/B of the prebuilt variants (tools/ab.py), then the GPU parity tests of the default build.
# usage: tools/gpu_ab.sh <tag> <variant names...>
TAG=${1:-ab}; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 600 python tools/ab.py run "$@" 2>&1 | tee $OUT/${TAG}_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -6 $OUT/${TAG}_pytest.txt
