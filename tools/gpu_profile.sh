#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): launch list + one full capture of the top kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --dims ${DIMS:-256}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_forward -s 3 -c 1 -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:classify_kernel\|emit_ -s 9 -c 3 -o gpurun_out/prof_mc $CMD > gpurun_out/ncu_mc.log 2>&1
tail -2 gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
