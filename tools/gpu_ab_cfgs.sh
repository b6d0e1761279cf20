#!/bin/bash
# On the GPU box: A/B of prebuilt variants over several configurations (SSB_AB = scene,variant,width,height,spp).
# usage: tools/gpu_ab_cfgs.sh <tag> "<cfg> <cfg> ..." <variant names...>
TAG=$1; CFGS=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
for cfg in $CFGS; do echo "== $cfg"; SSB_AB=$cfg timeout 300 python tools/ab.py run "$@"; done 2>&1 | tee $OUT/${TAG}_ab.txt
