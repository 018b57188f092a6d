#!/bin/bash
set -x
mkdir -p gpurun_out
# full capture of the predictor + corrector, product (fast) mode, 256^3: launches are ghost x6, predictor, corrector per step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vl_sweep_kernel -s 6 -c 2 -o gpurun_out/prof_fast_r1 python bench.py --steps 1 --warmup 3 --size 256 --mode fast --no-cpu-baseline --no-e2e > gpurun_out/prof_fast.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vl_sweep_kernel -s 6 -c 2 -o gpurun_out/prof_strict_r1 python bench.py --steps 1 --warmup 3 --size 256 --mode strict --no-cpu-baseline --no-e2e > gpurun_out/prof_strict.log 2>&1
# launch list of the reference's own kernels at 256^3
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed_pipe_fp64.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_ref_r1.csv python bench.py --impl reference --steps 2 --warmup 1 --size 256 > gpurun_out/ncu_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed_pipe_fp64.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fast_r1.csv python bench.py --steps 2 --warmup 3 --size 256 --mode fast --no-cpu-baseline --no-e2e > gpurun_out/ncu_fast.log 2>&1
ls -la gpurun_out
