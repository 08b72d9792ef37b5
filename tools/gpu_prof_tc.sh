#!/bin/bash
# one ncu --set full capture of the fused decoder kernel (after a plain run of the same command)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --query-only --dims ${DIMS:-128}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc2_forward -s 3 -c 1 -f -o gpurun_out/${OUT:-prof_tc2} $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
