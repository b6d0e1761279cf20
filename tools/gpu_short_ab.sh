#!/bin/bash
# On the GPU box (short slot): A/B of prebuilt variants (frame checksum per variant), the newest GPU tests, and the
# Jakob-Hanika config with prebaked textures (end-to-end leg: bake on the copy stream).
# Usage: tools/gpu_short_ab.sh <tag> <variant names...>
set -u
TAG=${1:-fin2}; shift; OUT=gpurun_out; mkdir -p $OUT
timeout 150 python tools/ab.py run "$@" 2>&1 | tee $OUT/${TAG}_ab.txt
timeout 120 python -m pytest tests/test_zz_gpu_prebake_progressive.py "tests/test_gpu_options.py::test_nearest_spectrum_filter" \
    "tests/test_gpu_parity.py::test_gpu_matches_oracle_and_reference_fixture" -m gpu -q -x > $OUT/${TAG}_pytest.txt 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -6 $OUT/${TAG}_pytest.txt
timeout 60 python bench.py --no-cpu-baseline --scene plane-srgb --variant jh --width 1024 --height 1024 --spp 64 --steps 5 --warmup 3 --prebake > $OUT/${TAG}_c4_jh_prebake.json 2> $OUT/${TAG}_c4_jh_prebake.err
echo "c4 prebake rc=$? $(cut -c1-200 $OUT/${TAG}_c4_jh_prebake.json)"; grep -o '"e2e": {[^}]*}' $OUT/${TAG}_c4_jh_prebake.json
