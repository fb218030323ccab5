#!/bin/bash
# 2-GPU call: full GPU test suite (incl. multi-GPU slab tests) + bench at N=1,2 in both exchange modes
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu2.log
tail -15 gpurun_out/pytest_gpu2.log
python bench.py --no-cpu-baseline > gpurun_out/bench2_n1.json 2> gpurun_out/bench2_n1.err; cat gpurun_out/bench2_n1.json | cut -c1-400
for mode in p2p nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --exchange $mode > gpurun_out/bench2_n2_$mode.json 2> gpurun_out/bench2_n2_$mode.err; echo "rc=$?"; cat gpurun_out/bench2_n2_$mode.json; tail -5 gpurun_out/bench2_n2_$mode.err
done
