#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 tools/slab_probe2.py z2z:1024 z2z:1024:FFTB200_SLAB_FUSED=0 c2c:1024 c2c:1024:FFTB200_SLAB_FUSED=0 z2z:512 d2z:1024 > gpurun_out/r02_slab_probe_exonly_n${N}.jsonl 2> gpurun_out/r02_slab_probe_exonly_n${N}.err; echo "probe rc=$?"
cat gpurun_out/r02_slab_probe_exonly_n${N}.jsonl; tail -n 3 gpurun_out/r02_slab_probe_exonly_n${N}.err
