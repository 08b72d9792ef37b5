#!/bin/bash
# round-2 ncu evidence (B200_PROFILING.md recipe): launch list of the headline part of the bench command, one full
# capture of the split-operand forward kernel at 256^3 and one of the split-operand latent step (8 shapes).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --latent-shapes 0 --slab-dims 0"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench256.csv $CMD > gpurun_out/ncu_launches.log 2>&1
F="python tools/prof_target.py fwd fp16x2"
$F > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc2_forward -s 2 -c 1 -f -o gpurun_out/r02_prof_tc2x2_256 $F > gpurun_out/ncu_full.log 2>&1
L="python tools/prof_target.py lat fp16x2"
$L > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:latent_tc2 -s 2 -c 1 -f -o gpurun_out/r02_prof_latent_x2 $L > gpurun_out/ncu_lat.log 2>&1
tail -2 gpurun_out/prof_plain.log | cut -c1-300; tail -2 gpurun_out/ncu_full.log; tail -2 gpurun_out/ncu_lat.log; ls -la gpurun_out/*.ncu-rep
