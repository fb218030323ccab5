#!/bin/bash
# final state on two GPUs: multi-GPU tests, bench at N=2 (with and without the compiled FFTW: the sampled-bin fall-back)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab.py -m gpu -q > gpurun_out/r02_pytest_slab_2gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_slab_2gpu_final.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err; echo "bench rc=$?"
FFTB200_BENCH_NO_FFTW=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 5 --warmup 3 --no-1024 > gpurun_out/r02_bench_n2_nofftw.json 2> gpurun_out/r02_bench_n2_nofftw.err; echo "bench(no fftw) rc=$?"
FFTB200_BENCH_NO_FFTW=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-1024 --no-cpu-baseline > gpurun_out/r02_bench_n1_nofftw.json 2> gpurun_out/r02_bench_n1_nofftw.err; echo "bench N=1 (no fftw) rc=$?"
python - <<PY
import json
for f in ("r02_bench_n2_final", "r02_bench_n2_nofftw", "r02_bench_n1_nofftw"):
    d=json.loads(open("gpurun_out/%s.json" % f).read().strip().split("\n")[-1])
    print(f, round(d["ms_per_step"],4), d["parity"], (d.get("scaling_1024") or {}).get("ms"))
PY
wc -l gpurun_out/r02_bench_n2_final.json
