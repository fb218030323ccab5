#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_slab.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu3.log
tail -15 gpurun_out/pytest_gpu3.log
for n in 512 1024; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 tools/slab_probe.py $n z2z 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee -a gpurun_out/slab_probe_n2.log
done
