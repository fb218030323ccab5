#!/bin/bash
# two GPUs: multi-GPU parity tests and the bench line at N=2 (parity gather, 1024^3 slab block, pipelined e2e)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab.py -m gpu -x -q > gpurun_out/r02_pytest_slab_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_slab_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_a.json 2> gpurun_out/r02_bench_n2_a.err; echo "bench rc=$?"
cut -c1-2500 gpurun_out/r02_bench_n2_a.json; tail -n 8 gpurun_out/r02_bench_n2_a.err
