#!/bin/bash
# round 2 evidence on ONE GPU: full GPU test suite, smoke, bench (both arms), ncu launch list, ncu full capture of the 512^3 passes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>/dev/null; cut -c1-250 gpurun_out/r02_bench_ref.json
python tools/cufft_compare.py > gpurun_out/r02_cufft_compare.jsonl 2> gpurun_out/r02_cufft_compare.err; cut -c1-170 gpurun_out/r02_cufft_compare.jsonl
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-1024 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-1024 > gpurun_out/ncu_launches.log 2>&1
python tools/prof_case.py z2z 512 512 512 > gpurun_out/r02_prof_case_z2z512.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fft_tile -s 3 -c 3 -o gpurun_out/r02_prof_z2z512 -f python tools/prof_case.py z2z 512 512 512 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log; ls -la gpurun_out/r02_prof_z2z512.ncu-rep
