#!/bin/bash
# one GPU: full GPU test suite, smoke, one default bench line (regression check after plan-builder changes)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_check.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r02_pytest_check.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_check.json 2> gpurun_out/r02_bench_check.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench_check.json
