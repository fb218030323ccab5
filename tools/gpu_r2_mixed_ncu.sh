#!/bin/bash
mkdir -p gpurun_out
set -x
REPS=2 python tools/mixed_one.py z2z 16384 1000 > gpurun_out/mixed_one_a.log 2>&1 || exit 1
REPS=2 python tools/mixed_one.py z2z 1 384 384 384 > gpurun_out/mixed_one_b.log 2>&1 || exit 1
REPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_mixed -c 2 -o gpurun_out/r02_mixed_1000 -f python tools/mixed_one.py z2z 16384 1000 > gpurun_out/ncu_a.log 2>&1
REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_mixed -c 3 -o gpurun_out/r02_mixed_384 -f python tools/mixed_one.py z2z 1 384 384 384 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
