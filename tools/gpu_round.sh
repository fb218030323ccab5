#!/bin/bash
# One gpurun call: GPU tests, bench, ncu launch list, ncu full capture of the 512^3 passes.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/prof_case.py z2z 512 512 512 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fft_tile -s 3 -c 3 -o gpurun_out/prof_z2z512 -f python tools/prof_case.py z2z 512 512 512 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
