#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline | tee gpurun_out/bench_quick.json
FFTB200_NO_FUSE=1 python bench.py --no-cpu-baseline | tee gpurun_out/bench_quick_nofuse.json | cut -c1-300
python tools/cufft_compare.py 2>&1 | tee gpurun_out/cufft_compare4.log | cut -c1-400
