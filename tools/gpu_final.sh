#!/bin/bash
# final-build evidence on ONE GPU: smoke, bench, reference arm, ncu launch list, ncu full capture of the 512^3 passes
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo "bench rc=$?"; cut -c1-220 gpurun_out/final_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2>/dev/null; cut -c1-220 gpurun_out/final_ref.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/prof_case.py z2z 512 512 512 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fft_tile -s 3 -c 3 -o gpurun_out/prof_z2z512_final -f python tools/prof_case.py z2z 512 512 512 > gpurun_out/ncu_full.log 2>&1
python tools/cufft_compare.py > gpurun_out/cufft_compare_final.log 2>&1; cut -c1-160 gpurun_out/cufft_compare_final.log
