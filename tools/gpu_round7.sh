#!/bin/bash
# 4-GPU validation of the final build: all GPU tests (slab tests at world 4), bench at N = 1, 2, 4
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu7.log
python bench.py --gpus 1 > gpurun_out/v4_n1.json 2> gpurun_out/v4_n1.err; cut -c1-200 gpurun_out/v4_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/v4_ref.json 2>/dev/null; cut -c1-200 gpurun_out/v4_ref.json
for n in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n > gpurun_out/v4_n$n.json 2> gpurun_out/v4_n$n.err; echo "rc=$?"; grep "^{" gpurun_out/v4_n$n.json | cut -c1-200
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29570 tools/slab_probe.py 1024 z2z one 2>&1 | grep "^{" | tee gpurun_out/slab_probe_1024_4gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 tools/slab_probe.py 1024 d2z quick 2>&1 | grep "^{" | tee -a gpurun_out/slab_probe_1024_4gpu.log
