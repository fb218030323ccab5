#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29570 tools/slab_probe.py 1024 d2z quick 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee gpurun_out/slab_probe_1024_d2z_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/slab_probe.py 512 z2z quick 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee -a gpurun_out/slab_probe_1024_d2z_2gpu.log
