#!/bin/bash
mkdir -p gpurun_out
for c in "z2z 1024 1024 64" "c2c 512 512 512" "d2z 4096 4096" "c2c 134217728"; do
  tag=$(echo $c | tr ' ' '_')
  python tools/prof_case.py $c > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:fft_ -s 3 -c 3 -o gpurun_out/prof_$tag -f python tools/prof_case.py $c > gpurun_out/ncu_$tag.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
