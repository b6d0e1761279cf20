#!/bin/bash
# On the GPU box: bench lines for the other SURVEY 8(d) configurations (not the headline): C3, C4/C5 at reduced spp, RGB mode.
TAG=${1:-cfg}; OUT=gpurun_out; mkdir -p $OUT
run() { name=$1; shift; timeout 900 python bench.py --no-cpu-baseline "$@" > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err; echo "$name rc=$? $(cut -c1-160 $OUT/${TAG}_${name}.json)"; }
run c3_cornell_2006_spp256 --scene cornell --variant ours2006 --spp 256 --steps 5 --warmup 3
run c4_plane_jh_1024_spp64 --scene plane-srgb --variant jh --width 1024 --height 1024 --spp 64 --steps 5 --warmup 3
run c5_cornellsrgb_meng_2048_spp16 --scene cornell-srgb --variant meng --width 2048 --height 2048 --spp 16 --steps 5 --warmup 3
run rgb_cornellsrgb_512_spp64 --scene cornell-srgb --variant rgb --steps 10 --warmup 3
