#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): launch list of one bench run, one full
# capture of the fused decoder kernel at the bench size, one of the marching-cubes kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --latent-shapes 0 --dims ${DIMS:-256}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc2_forward -s 3 -c 1 -f -o gpurun_out/prof_tc2_256 $CMD > gpurun_out/ncu_full.log 2>&1
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:classify_kernel\|emit_\|scan_groups\|scan_totals -s 15 -c 5 -f -o gpurun_out/prof_mc $CMD > gpurun_out/ncu_mc.log 2>&1
tail -2 gpurun_out/prof_plain.log | cut -c1-600; tail -3 gpurun_out/ncu_full.log; tail -3 gpurun_out/ncu_mc.log
