#!/bin/bash
# On the GPU box: launch list (time, DRAM bytes, instruction counts per launch) of one bench run + one full ncu capture of
# the depth-1 intersect / sort / shade launches.   usage: tools/gpu_prof2.sh <tag>
TAG=${1:-prof}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_list_summary.py $OUT/${TAG}_launches.csv $OUT/${TAG}_trace_kernel_traffic.json > $OUT/${TAG}_launch_list_summary.txt 2>&1
tail -22 $OUT/${TAG}_launch_list_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ssb_intersect|ssb_shade|ssb_bin|ssb_fold|ssb_accumulate" -s 80 -c 4 -f -o $OUT/${TAG}_trace \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT | tail -8
