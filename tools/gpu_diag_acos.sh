#!/bin/bash
# On the GPU box: tools/diag_acos_pairs.py on the current build, the rest of tests/test_gpu_parity.py, and the same diagnosis on a
# prebuilt scalar-acosf variant (variants/libssb200_pairs0.so from `tools/ab.py build pairs0=SSB_ACOS_PAIRS=0`).
# This is the call that showed the r4d stream-ordering race (profiles/r4d_diag_eval_math_race.txt).
set -u
TAG=${1:-fin4}; OUT=gpurun_out; mkdir -p $OUT
timeout 40 python tools/diag_acos_pairs.py > $OUT/${TAG}_diag_cur.txt 2>&1; echo "diag rc=$?"; cat $OUT/${TAG}_diag_cur.txt | cut -c1-300
timeout 80 python -m pytest tests/test_gpu_parity.py -m gpu -q --durations=6 \
   --deselect tests/test_gpu_parity.py::test_device_paired_acosf_equals_scalar_exhaustively \
   --deselect tests/test_gpu_parity.py::test_gpu_matches_oracle_and_reference_fixture \
   --deselect tests/test_gpu_parity.py::test_device_math_bit_exact > $OUT/${TAG}_pytest.txt 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -12 $OUT/${TAG}_pytest.txt
cp variants/libssb200_pairs0.so simple-spectral_b200/libssb200.so
timeout 30 python tools/diag_acos_pairs.py > $OUT/${TAG}_diag_pairs0.txt 2>&1; echo "diag pairs0 rc=$?"; head -8 $OUT/${TAG}_diag_pairs0.txt | cut -c1-200
