#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 8 > gpurun_out/v5_n8.json 2> gpurun_out/v5_n8.err; echo "rc=$?"; grep "^{" gpurun_out/v5_n8.json | cut -c1-260
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29570 tools/slab_probe.py 1024 z2z one 2>&1 | grep "^{" | tee gpurun_out/slab_probe_1024_8gpu_v5.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/slab_probe.py 1024 d2z quick 2>&1 | grep "^{" | tee -a gpurun_out/slab_probe_1024_8gpu_v5.log
python tools/cufft_compare.py 2>&1 | grep -E "1024, 1024, 1024|512, 512, 512" | cut -c1-200 | tee gpurun_out/single_on_8gpu_box.log
