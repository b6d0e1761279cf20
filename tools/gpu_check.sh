#!/bin/bash
# Run on the GPU box (via gpurun): GPU tests, launch list (time, DRAM bytes, instruction counts per launch), bench,
# and one full ncu capture of a depth-1 intersect/sort/shade sequence.
# Usage: tools/gpu_check.sh <tag> [noprof]      outputs -> gpurun_out/<tag>_*
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt
tail -5 $OUT/${TAG}_pytest.txt
if [ "${2:-}" != "noprof" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1; echo "ncu launches rc=$?"
# the bench line below reads the per-frame DRAM bytes / instruction counts of THIS build from the summary
python tools/launch_list_summary.py $OUT/${TAG}_launches.csv profiles/trace_kernel_traffic.json > $OUT/${TAG}_launch_list_summary.txt 2>&1 && cp profiles/trace_kernel_traffic.json $OUT/${TAG}_trace_kernel_traffic.json
tail -22 $OUT/${TAG}_launch_list_summary.txt
fi
timeout 600 python bench.py --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
if [ "${2:-}" != "noprof" ]; then
# one whole frame of bounce launches (depths 0..8) + its finalize, after two warm frames
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"ssb_intersect|ssb_shade|ssb_bin|ssb_fold|ssb_accumulate" -s 80 -c 4 -f -o $OUT/${TAG}_trace \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT | tail -12
fi
