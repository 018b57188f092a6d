#!/bin/bash
mkdir -p gpurun_out
MHD_ARM=fast timeout 900 ncu --set full --clock-control none --import-source on -k regex:mhd_dir_kernel -s 18 -c 6 -o gpurun_out/prof_r2_mhd_fast_256 python bench_mhd.py 256 2 > gpurun_out/prof_r2_mhd.log 2>&1
ls -la gpurun_out/*.ncu-rep
