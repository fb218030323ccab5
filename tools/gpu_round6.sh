#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu6.log
python bench.py > gpurun_out/bench_n1_v3.json 2> gpurun_out/bench_n1_v3.err; cut -c1-250 gpurun_out/bench_n1_v3.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft -c 400 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/prof_case.py z2z 512 512 512 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fft_tile -s 3 -c 3 -o gpurun_out/prof_z2z512_v3 -f python tools/prof_case.py z2z 512 512 512 > gpurun_out/ncu_full.log 2>&1
python tools/cufft_compare.py > gpurun_out/cufft_compare7.log 2>&1; cut -c1-200 gpurun_out/cufft_compare7.log
python tools/sanitize_case.py > gpurun_out/plain3.log 2>&1 &&
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -5 gpurun_out/sanitizer_memcheck.log
