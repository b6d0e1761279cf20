#!/bin/bash
# On the GPU box: the whole GPU suite of the final build, then the headline bench line.
set -u
TAG=${1:-fin3}; OUT=gpurun_out; mkdir -p $OUT
timeout 170 python -m pytest tests -m gpu -q -x --durations=8 > $OUT/${TAG}_pytest.txt 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -14 $OUT/${TAG}_pytest.txt
timeout 40 python bench.py --steps 25 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $OUT/${TAG}_bench.json
