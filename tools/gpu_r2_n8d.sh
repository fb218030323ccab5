#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/slab_probe2.py d2z:1024 d2z:1024:FFTB200_SLAB_COL_CHUNKS=4 d2z:1024:FFTB200_SLAB_COL_CHUNKS=4,FFTB200_SLAB_EX_CTAS=120 d2z:1024:FFTB200_SLAB_COL_CHUNKS=4,FFTB200_SLAB_EX_CTAS=74 d2z:1024:FFTB200_SLAB_COL_CHUNKS=2 d2z:1024:FFTB200_SLAB_COL_CHUNKS=8 d2z:512 d2z:512:FFTB200_SLAB_COL_CHUNKS=1 d2z:512:FFTB200_SLAB_COL_CHUNKS=2 > gpurun_out/r02_slab_probe_d2z_n8.jsonl 2> gpurun_out/r02_slab_probe_d2z_n8.err; echo "probe rc=$?"
cat gpurun_out/r02_slab_probe_d2z_n8.jsonl; tail -n 3 gpurun_out/r02_slab_probe_d2z_n8.err
