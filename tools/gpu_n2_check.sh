#!/bin/bash
# On a 2-GPU box: the driver's own N=2 bench launch, the in-library multi-device tests on real peers, the newest tests.
TAG=${1:-n2}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-300 $OUT/${TAG}_bench_n2.json; grep -o '"strong": {.*' $OUT/${TAG}_bench_n2.json | cut -c1-700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/${TAG}_ref_n2.json 2> $OUT/${TAG}_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-200 $OUT/${TAG}_ref_n2.json
timeout 600 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_render_fuzz.py -m gpu -q -k "multidevice or extreme or multi" > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -4 $OUT/${TAG}_pytest.txt
