#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29691 tools/slab_probe2.py z2z:1024:FFTB200_SLAB_FUSED=0 z2z:1024:FFTB200_SLAB_EX_CTAS=98 z2z:1024:FFTB200_SLAB_EX_CTAS=120 z2z:1024:FFTB200_SLAB_EX_CTAS=74 z2z:1024:FFTB200_SLAB_EX_CTAS=98,FFTB200_SLAB_PLANE_CHUNKS=8 z2z:1024:FFTB200_SLAB_EX_CTAS=98,FFTB200_SLAB_PLANE_CHUNKS=2 c2c:1024:FFTB200_SLAB_EX_CTAS=98 c2c:1024:FFTB200_SLAB_FUSED=0 > gpurun_out/r02_slab_probe_exonly_n8.jsonl 2> gpurun_out/r02_slab_probe_exonly_n8.err; echo "probe rc=$?"
cat gpurun_out/r02_slab_probe_exonly_n8.jsonl; tail -n 3 gpurun_out/r02_slab_probe_exonly_n8.err
