#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/cufft_compare.py 2>&1 | tee gpurun_out/cufft_compare5.log | cut -c1-700
python bench.py | tee gpurun_out/bench_n1_v2.json | cut -c1-200
