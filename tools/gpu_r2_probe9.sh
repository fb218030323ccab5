#!/bin/bash
# ragged block inside the main launch (D2Z), fp32 four-step twiddle step in fp32, 512-point row alternates
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not multi_gpu and not 1024cubed_z2z" > gpurun_out/r02_pytest_i.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_i.log
A=gpurun_out/r02_alt_probe9.jsonl; : > $A
E=gpurun_out/r02_alt_probe9.err
timeout 600 python tools/alt_probe.py d2z 1024,1024,1024 0:0,1:0,2:0,3:0,4:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py d2z 512,512,512 0:0,1:0,2:0,3:0,4:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 512,512,512 0:0,1:0,2:0,3:0,4:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 134217728 0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 16777216 0 >> $A 2>> $E
python - <<PY
import json
for l in open("$A"):
    d=json.loads(l); print(d["kind"], d["shape"], d["alt"], d.get("ms"), d.get("rel_l2_vs_alt0"), d.get("error"))
    for p in d.get("passes",[]): print("    ", p[:120])
PY
tail -n 5 $E
