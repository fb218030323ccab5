#!/bin/bash
mkdir -p gpurun_out
python tools/prof_case.py z2z 512 512 512 > gpurun_out/plain_f.log 2>&1 || exit 1
for cfg in "128 2" "64 2" "64 1" "128 1" "256 2" "512 3"; do
set -- $cfg
echo "== tiles=$1 lag=$2"
FFTB200_FUSE_TILES=$1 FFTB200_FUSE_LAG=$2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none -k regex:fft_ -s 2 -c 2 python tools/prof_case.py z2z 512 512 512 2>&1 | grep -E "fft_|duration|dram__bytes|hit_rate|srcunit"
done 2>&1 | tee gpurun_out/fuse_probe.log
