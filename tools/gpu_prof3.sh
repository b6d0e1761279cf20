#!/bin/bash
# On the GPU box: one full ncu capture of the depth-1 intersect / sort / shade launches of the CURRENT library.
TAG=${1:-prof}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ssb_intersect|ssb_shade|ssb_bin|ssb_fold|ssb_accumulate" -s 80 -c 4 -f -o $OUT/${TAG}_trace \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
