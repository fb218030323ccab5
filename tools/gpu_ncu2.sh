#!/bin/bash
mkdir -p gpurun_out
python tools/prof_case.py z2z 1024 1024 64 > gpurun_out/plain_a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fft_ -s 3 -c 3 -o gpurun_out/prof_z2z_1024x1024x64 -f python tools/prof_case.py z2z 1024 1024 64 > gpurun_out/ncu_a.log 2>&1
python tools/prof_case.py c2c 512 512 512 > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fft_ -s 3 -c 3 -o gpurun_out/prof_c2c_512 -f python tools/prof_case.py c2c 512 512 512 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
