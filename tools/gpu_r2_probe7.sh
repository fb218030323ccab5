#!/bin/bash
# round 2: fused slab kernel on one GPU (G=1), real-transform blocked layout, fp32 four-step twiddle, persistent clusters
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not multi_gpu and not 1024cubed" > gpurun_out/r02_pytest_g.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_g.log
A=gpurun_out/r02_alt_probe7.jsonl; : > $A
E=gpurun_out/r02_alt_probe7.err
for c in "d2z 1024,1024,1024" "d2z 512,512,512" "c2c 134217728" "z2z 16777216"; do
  timeout 600 python tools/alt_probe.py $c 0 >> $A 2>> $E
done
timeout 600 python tools/alt_probe.py d2z 1024,1024,1024 0 FFTB200_ZBLOCK=0 >> $A 2>> $E
for c in "z2z 4096,4096" "d2z 4096,4096" "z2z 8192,8192" "c2c 4096,4096"; do
  timeout 600 python tools/alt_probe.py $c 0 FFTB200_CLUSTER_PERSIST=1 >> $A 2>> $E
done
cut -c1-800 $A
tail -n 5 $E
timeout 600 python -m pytest tests -m gpu -x -q -k "c5_1024cubed" > gpurun_out/r02_pytest_g2.log 2>&1; echo "pytest-d2z-1024 rc=$?"; tail -3 gpurun_out/r02_pytest_g2.log
