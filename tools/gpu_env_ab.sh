#!/bin/bash
# On the GPU box: A/B of run-time knobs (environment variables) on the current library: frame time + checksum per setting.
# usage: tools/gpu_env_ab.sh <tag> "VAR=val" "VAR=val" ...   ("-" = no variable)
TAG=$1; shift; OUT=gpurun_out; mkdir -p $OUT
cp simple-spectral_b200/libssb200.so variants/libssb200_envab.so
for kv in "$@"; do
  if [ "$kv" = "-" ]; then echo "== default" | tee -a $OUT/${TAG}_env_ab.txt; timeout 120 python tools/ab.py run envab 2>&1 | tee -a $OUT/${TAG}_env_ab.txt
  else echo "== $kv" | tee -a $OUT/${TAG}_env_ab.txt; env "$kv" timeout 120 python tools/ab.py run envab 2>&1 | tee -a $OUT/${TAG}_env_ab.txt; fi
done
