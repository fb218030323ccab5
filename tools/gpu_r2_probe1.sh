#!/bin/bash
# round 2, first GPU call: data-movement ceilings (LDG vs TMA, near/far strides) and alternative 1024-point tile rows
mkdir -p gpurun_out
nvidia-smi -L | head -2
O=gpurun_out/r02_tma_probe.jsonl; : > $O
for cfg in "512 512 512 1" "512 512 512 0" "1024 1024 1024 1" "1024 1024 1024 0"; do
  timeout 120 tools/bin/tma_probe $cfg 5 >> $O 2>> gpurun_out/r02_tma_probe.err || echo "{\"error\": \"tma_probe $cfg rc=$?\"}" >> $O
done
cat $O
A=gpurun_out/r02_alt_probe.jsonl; : > $A
timeout 600 python tools/alt_probe.py z2z 1024,1024,1024 0,1,2,3 >> $A 2>> gpurun_out/r02_alt_probe.err
timeout 600 python tools/alt_probe.py z2z 1024,1024,1024 0,2,3 FFTB200_ZBLOCK=0 >> $A 2>> gpurun_out/r02_alt_probe.err
timeout 600 python tools/alt_probe.py z2z 1024,1024,1024 2,3 FFTB200_PREFETCH=0 >> $A 2>> gpurun_out/r02_alt_probe.err
timeout 600 python tools/alt_probe.py c2c 1024,1024,1024 0,1,2 >> $A 2>> gpurun_out/r02_alt_probe.err
timeout 600 python tools/alt_probe.py d2z 1024,1024,1024 0,2,3 >> $A 2>> gpurun_out/r02_alt_probe.err
timeout 600 python tools/alt_probe.py z2z 1024,1024,64 0,1,2,3 >> $A 2>> gpurun_out/r02_alt_probe.err
cat $A | cut -c1-700
tail -5 gpurun_out/r02_alt_probe.err gpurun_out/r02_tma_probe.err
