#!/bin/bash
# usage: gpu_prof3.sh fast|strict   (one report per call: two exceed the 64 MiB return limit)
mkdir -p gpurun_out
m=$1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vl_dir_kernel -s 18 -c 6 -o gpurun_out/prof_r2_${m}_512 python bench.py --steps 1 --warmup 3 --size 512 --mode $m --no-cpu-baseline --no-e2e > gpurun_out/prof_r2_$m.log 2>&1
ls -la gpurun_out/*.ncu-rep
