#!/bin/bash
# 1-GPU call: the whole GPU test suite on the new arithmetic, ncu of one strict step, 256^3 numbers
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest_all.log
tail -4 gpurun_out/r2_pytest_all.log
run() { tag=$1; shift
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --single-mode "$@" 2>gpurun_out/r2_bench_$tag.err | tee gpurun_out/r2_bench_$tag.json | python tools/short.py $tag | tee -a gpurun_out/r2_run3.txt; }
run kh256_strict --mode strict --size 256 --workload kh
run kh256_fast --mode fast --size 256 --workload kh
# ncu: warm-up 3 steps, then one step's six sweeps (per step: 6 sweeps + 6 ghost fills; initial calc_dti + 6 ghost fills)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:vl_dir_kernel -s 18 -c 6 -f -o gpurun_out/r3_prof_strict_512 python bench.py --steps 1 --warmup 3 --mode strict --single-mode --no-cpu-baseline --no-e2e > gpurun_out/r3_prof_strict.log 2>&1
ls -la gpurun_out/*.ncu-rep
