#!/bin/bash
# On the GPU box: parity tests, then a knob sweep (rebuilds on the box), then the bench.
TAG=${1:-tune}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -4 $OUT/${TAG}_pytest.txt
timeout 1500 python tools/tune.py "$@" 2>&1 | tee $OUT/${TAG}_tune.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json | cut -c1-400
