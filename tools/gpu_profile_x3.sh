#!/bin/bash
# ncu full capture of the three-pass split-operand forward at 256^3 (after the same command ran clean without ncu)
mkdir -p gpurun_out
F="python tools/prof_target.py fwd fp16x2"
timeout 200 python tools/x2_check.py 2>&1 | grep "fp16x2\|256^3" | tail -10
$F > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc2_forward -s 2 -c 1 -f -o gpurun_out/r02_prof_tc2x3_256 $F > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
