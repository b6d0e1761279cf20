#!/bin/bash
# On the GPU box: A/B of prebuilt variants on the headline frame AND on a strong-scaling slice of it (512x64: what one of
# eight GPUs renders), then the GPU test suite of the default build.
# usage: tools/gpu_ab2.sh <tag> <variant names...>
TAG=${1:-ab}; shift
OUT=gpurun_out; mkdir -p $OUT
{
timeout 400 python tools/ab.py run "$@"
echo "SLICE 512x64 spp64"
SSB_AB="cornell-srgb,ours1931,512,64,64" timeout 300 python tools/ab.py run "$@"
} 2>&1 | tee $OUT/${TAG}_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -6 $OUT/${TAG}_pytest.txt
