#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/cufft_compare.py 2>&1 | grep -E '"c2c"|16777216' | cut -c1-650
