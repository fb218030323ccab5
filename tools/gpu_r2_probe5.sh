#!/bin/bash
# round 2, fifth GPU call: narrow-tile copy ceilings, W=4 alternates for 1024-point columns, conj-based inverse parity
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not multi_gpu and not 1024cubed" > gpurun_out/r02_pytest_e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_e.log
O=gpurun_out/r02_tma_probe2.jsonl; : > $O
for cfg in "1024 1024 1024 1" "1024 1024 1024 0"; do
  timeout 120 tools/bin/tma_probe $cfg 5 >> $O 2>> gpurun_out/r02_tma_probe2.err || echo "{\"error\": \"tma_probe $cfg rc=$?\"}" >> $O
done
grep -E "flat|ldg" $O
A=gpurun_out/r02_alt_probe5.jsonl; : > $A
E=gpurun_out/r02_alt_probe5.err
timeout 600 python tools/alt_probe.py z2z 1024,1024,1024 0:0,0:2 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 1024,1024,1024 0:2 FFTB200_ZBLOCK=0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 1024,1024,1024 0:0,0:2 >> $A 2>> $E
timeout 600 python tools/alt_probe.py d2z 1024,1024,1024 0:0,0:2 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 4096,4096 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py d2z 4096,4096 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 512,512,512 0:0 >> $A 2>> $E
cut -c1-800 $A
tail -n 5 $E
