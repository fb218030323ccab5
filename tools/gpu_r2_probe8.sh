#!/bin/bash
# cluster-kernel alternates for 4096 / 8192-point columns; R=8 x 1024-thread alternate for 1024-point columns; fixed tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "error_codes or bluestein or inplace_r2c or normalisation" > gpurun_out/r02_pytest_h.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_h.log
A=gpurun_out/r02_alt_probe8.jsonl; : > $A
E=gpurun_out/r02_alt_probe8.err
timeout 600 python tools/alt_probe.py z2z 4096,4096 0:0,0:2,0:3,0:4 >> $A 2>> $E
timeout 600 python tools/alt_probe.py d2z 4096,4096 0:0,0:2,0:3,0:4 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 8192,8192 0:0,0:1 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 1024,1024,1024 0:0,0:2 >> $A 2>> $E
python - <<PY
import json
for l in open("$A"):
    d=json.loads(l); print(d["kind"], d["shape"], d["alt"], d.get("ms"), d.get("rel_l2_vs_alt0"), d.get("error"))
    for p in d.get("passes",[]): print("    ", p[:120])
PY
tail -n 5 $E
