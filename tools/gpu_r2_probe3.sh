#!/bin/bash
# round 2, third GPU call: parity after the twiddle-table change, row-alternative sweep, bench.py with the parity / 1024 blocks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not multi_gpu and not 1024cubed" > gpurun_out/r02_pytest_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_c.log
A=gpurun_out/r02_alt_probe3.jsonl; : > $A
E=gpurun_out/r02_alt_probe3.err
timeout 600 python tools/alt_probe.py z2z 1024,1024,1024 0:0,1:0,2:0,3:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 4096,4096 0:0,1:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py d2z 4096,4096 0:0,1:0,2:0,3:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 2048,2048 0:0,1:0,2:0,3:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 512,512,512 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 1024,1024,1024 0:0,1:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py c2c 134217728 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py z2z 8192,8192 0:0 >> $A 2>> $E
timeout 600 python tools/alt_probe.py d2z 1024,1024,1024 0:0 >> $A 2>> $E
cut -c1-800 $A
tail -n 5 $E
timeout 900 python bench.py > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/r02_bench_n1_a.json; tail -n 5 gpurun_out/r02_bench_n1_a.err
