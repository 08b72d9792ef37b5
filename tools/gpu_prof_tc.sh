#!/bin/bash
mkdir -p gpurun_out
export DEEPJOIN_B200_CLUSTER=${C:-1}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --dims 128"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_forward -s 3 -c 1 -f -o gpurun_out/prof_tc_v2 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
