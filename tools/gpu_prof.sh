#!/bin/bash
# On the GPU box: (optional knob sweep) + full ncu capture of one frame's bounce/finalize launches.
TAG=${1:-prof}; shift
OUT=gpurun_out; mkdir -p $OUT
if [ $# -gt 0 ]; then timeout 1500 python tools/tune.py "$@" 2>&1 | tee $OUT/${TAG}_tune.txt; fi
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"ssb_intersect|ssb_shade|ssb_bin|ssb_fold|ssb_accumulate" -s 80 -c 4 -f -o $OUT/${TAG}_trace \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
