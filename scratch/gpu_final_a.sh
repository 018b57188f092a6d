#!/bin/bash
# usage: gpu_final_a.sh fast|strict
mkdir -p gpurun_out
m=$1
if [ $m = strict ]; then
timeout 600 python -m pytest tests/test_gpu_exact_div.py -q -x 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_512.csv python bench.py --steps 2 --warmup 3 --single-mode --no-cpu-baseline --no-e2e > gpurun_out/r2_launches_bench.log 2>&1
fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vl_dir_kernel -s 18 -c 6 -o gpurun_out/prof_r2_${m}_512 python bench.py --steps 1 --warmup 3 --size 512 --mode $m --single-mode --no-cpu-baseline --no-e2e > gpurun_out/prof_r2_$m.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
