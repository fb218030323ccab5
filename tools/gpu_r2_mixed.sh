#!/bin/bash
# one GPU: the mixed-radix tests, then the probe against the generic path and cuFFT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed or bluestein or golden or c2r or inplace or in_place" > gpurun_out/mixed_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/mixed_pytest.log
tail -5 gpurun_out/mixed_pytest.log
timeout 400 python tools/mixed_probe.py > gpurun_out/mixed_probe.jsonl 2> gpurun_out/mixed_probe.err
echo "probe exit $?"
cut -c1-900 gpurun_out/mixed_probe.jsonl
tail -3 gpurun_out/mixed_probe.err
