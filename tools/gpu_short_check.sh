#!/bin/bash
# On the GPU box, for a SHORT slot (a few minutes): the newest GPU tests + a representative slice of the parity suites,
# the headline bench line, and the Jakob-Hanika config with and without prebaked coefficient textures.  Most important first.
# Usage: tools/gpu_short_check.sh <tag>      outputs -> gpurun_out/<tag>_*
set -u
TAG=${1:-fin}; OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_zz_gpu_prebake_progressive.py tests/test_gpu_options.py \
    "tests/test_gpu_parity.py::test_gpu_matches_oracle_and_reference_fixture" tests/test_cli.py -m gpu -q -x --durations=5 > $OUT/${TAG}_pytest.txt 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -12 $OUT/${TAG}_pytest.txt
timeout 90 python bench.py --steps 25 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json
for pb in "" "--prebake"; do
  timeout 60 python bench.py --no-cpu-baseline --scene plane-srgb --variant jh --width 1024 --height 1024 --spp 64 --steps 5 --warmup 3 $pb > $OUT/${TAG}_c4_jh${pb#-}.json 2> $OUT/${TAG}_c4_jh${pb#-}.err
  echo "c4 $pb rc=$? $(cut -c1-200 $OUT/${TAG}_c4_jh${pb#-}.json)"
done
