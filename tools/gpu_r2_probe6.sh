#!/bin/bash
# round 2, sixth GPU call: parity + timings after the instruction-count work (store addressing, fast division, conj inverse)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not multi_gpu and not 1024cubed" > gpurun_out/r02_pytest_f.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_f.log
A=gpurun_out/r02_alt_probe6.jsonl; : > $A
E=gpurun_out/r02_alt_probe6.err
for c in "z2z 512,512,512" "z2z 1024,1024,1024" "d2z 1024,1024,1024" "c2c 1024,1024,1024" "c2c 512,512,512" "z2z 4096,4096" "d2z 4096,4096" "c2c 134217728" "z2z 16777216" "z2z 256,256,256" "z2z 8192,8192" "c2c 4096,4096" "z2z 2048,2048"; do
  timeout 600 python tools/alt_probe.py $c 0 >> $A 2>> $E
done
cut -c1-800 $A
tail -n 5 $E
