#!/bin/bash
# On an N-GPU box: the driver's own bench launch at N ranks (weak value + strong section), and a BASELINE multi-GPU
# configuration at full size, row-band sharded.   usage: tools/gpu_n_final.sh <tag> <N> <bench args of the big config...>
TAG=$1; N=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 50 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench n$N rc=$?"; cut -c1-160 $OUT/${TAG}_bench_n$N.json; grep -o '"strong": {.*' $OUT/${TAG}_bench_n$N.json | cut -c1-420
timeout 900 $TR --master-port 29522 bench.py --gpus $N --no-cpu-baseline --shard tiles "$@" > $OUT/${TAG}_big_n$N.json 2> $OUT/${TAG}_big_n$N.err; echo "big n$N rc=$?"; cut -c1-200 $OUT/${TAG}_big_n$N.json; grep -o '"frame_sha": "[0-9a-f]*"' $OUT/${TAG}_big_n$N.json | head -2
