#!/bin/bash
# round 2, second GPU call: parity of the rebuilt library + alternative rows for long contiguous / strided axes
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q -k "not multi_gpu and not 1024cubed" --durations=8 > gpurun_out/r02_pytest_a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_a.log
A=gpurun_out/r02_alt_probe2.jsonl; : > $A
E=gpurun_out/r02_alt_probe2.err
timeout 600 python tools/alt_probe.py z2z 1024,1024,1024 0:0,1:0,2:0,3:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 4096,4096 0:0,1:0,0:1 >> $A 2>> $E
timeout 600 python tools/alt_probe.py d2z 4096,4096 0:0,1:0,2:0,3:0,0:1 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 2048,2048 0:0,1:0,2:0,3:0,0:1 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 1024,1024,1024 0:0,1:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 4096,4096 0:0,1:0,0:1 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 2048,2048 0:0,1:0,0:1 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 8192,8192 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py d2z 1024,1024,1024 0:0 >> $A 2>> $E
cut -c1-800 $A
tail -n 5 $E
timeout 900 python -m pytest tests -m gpu -x -q -k "1024cubed" --durations=4 > gpurun_out/r02_pytest_b.log 2>&1; echo "pytest-1024 rc=$?"; tail -8 gpurun_out/r02_pytest_b.log
