#!/bin/bash
# two GPUs: fused single-kernel slab path (two queues, two roles) vs the multi-launch path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab.py -m gpu -x -q -k "single_kernel or multi_gpu" > gpurun_out/r02_pytest_slab_2gpu_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_slab_2gpu_c.log
for fused in 1 0; do
FFTB200_SLAB_FUSED=$fused timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$fused bench.py --gpus 2 --steps 20 --warmup 5 --no-parity > gpurun_out/r02_bench_n2c_fused$fused.json 2> gpurun_out/r02_bench_n2c_fused$fused.err; echo "bench fused=$fused rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n2c_fused$fused.json").read().strip().split("\n")[-1])
print("fused=$fused", "512^3 ms", round(d["ms_per_step"],4), "launches/step", d["gpu_launches"]/d["steps"], "1024^3 ms", round(d["scaling_1024"]["ms"],3), d["scaling_1024"]["parity"]["ok"], "e2e GB/s", round(d["e2e"]["host_link_GB/s_each_way"],1))
PY
done
tail -n 4 gpurun_out/r02_bench_n2c_fused1.err
