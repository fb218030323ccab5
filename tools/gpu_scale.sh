#!/bin/bash
# scaling run: bench.py at N = 1, 2, 4, 8 (as the driver launches it) + slab tests on all GPUs + 1024^3 probe
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
python -m pytest tests/test_slab.py -m gpu -q > gpurun_out/pytest_slab_${NG}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_slab_${NG}gpu.log
python bench.py --gpus 1 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; cut -c1-330 gpurun_out/scale_n1.json
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "rc=$?"; cut -c1-330 gpurun_out/scale_n$n.json; tail -2 gpurun_out/scale_n$n.err
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --exchange nccl > gpurun_out/scale_nccl_n$n.json 2> gpurun_out/scale_nccl_n$n.err; cut -c1-330 gpurun_out/scale_nccl_n$n.json
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29570 tools/slab_probe.py 1024 z2z quick 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee gpurun_out/slab_probe_1024_${NG}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 tools/slab_probe.py 1024 d2z quick 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee -a gpurun_out/slab_probe_1024_${NG}gpu.log
