#!/bin/bash
# On the GPU box: everything the final build of a round is judged on, in one call —
#   A/B of the named variants (frame checksum), the GPU test suite, smoke(), the default bench line, the launch list + traffic
#   json + one full ncu capture of the depth-1 launches (tools/gpu_prof2.sh), and the other configurations (tools/gpu_configs.sh).
# usage: tools/gpu_final.sh <tag> [variant names...]
TAG=${1:-fin}; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ $# -gt 0 ]; then timeout 300 python tools/ab.py run "$@" 2>&1 | tee $OUT/${TAG}_ab.txt; fi
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -4 $OUT/${TAG}_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.txt 2>&1; tail -1 $OUT/${TAG}_smoke.txt
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-240 $OUT/${TAG}_bench.json
tools/gpu_prof2.sh $TAG
# the bench line again, now that the traffic json of THIS build exists (roofline.traffic / issue filled in)
cp $OUT/${TAG}_trace_kernel_traffic.json profiles/trace_kernel_traffic.json
timeout 300 python bench.py --no-cpu-baseline --steps 50 > $OUT/${TAG}_bench_with_traffic.json 2>> $OUT/${TAG}_bench.err; grep -o '"roofline": {[^}]*}' $OUT/${TAG}_bench_with_traffic.json | cut -c1-400
tools/gpu_configs.sh $TAG
