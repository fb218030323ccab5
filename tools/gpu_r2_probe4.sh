#!/bin/bash
# round 2, fourth GPU call: parity after the per-stage twiddle tables, sweep, ncu of the 1024-point column pass and 4096-point row pass
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not multi_gpu and not 1024cubed" > gpurun_out/r02_pytest_d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_d.log
A=gpurun_out/r02_alt_probe4.jsonl; : > $A
E=gpurun_out/r02_alt_probe4.err
timeout 600 python tools/alt_probe.py z2z 1024,1024,1024 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 4096,4096 0:0,1:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py d2z 4096,4096 0:0,1:0,2:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 2048,2048 0:0,1:0,2:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 512,512,512 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 512,512,512 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 1024,1024,1024 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 134217728 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 8192,8192 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 16777216 0:0 >> $A 2>> $E
cut -c1-800 $A
tail -n 5 $E
python tools/prof_case.py z2z 1024 1024 128 > gpurun_out/plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fft_tile -c 3 -f -o gpurun_out/r02_prof_z2z_1024_1024_128 python tools/prof_case.py z2z 1024 1024 128 > gpurun_out/ncu1.log 2>&1
python tools/prof_case.py z2z 4096 4096 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fft_ -c 2 -f -o gpurun_out/r02_prof_z2z_4096_4096 python tools/prof_case.py z2z 4096 4096 > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu1.log gpurun_out/ncu2.log
ls -la gpurun_out/*.ncu-rep
