#!/bin/bash
# dual energy / scalars / static gravity parity, drop-in builds, and the headline with the temperature floor on by default
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_passive.py tests/test_gpu_gravity.py tests/test_gpu_dropin.py -q 2>&1 | tail -60 ) 2>&1 | tee gpurun_out/r10_pytest_new.log
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity"
python bench.py $q 2>gpurun_out/r10_bench_strict.err | tee gpurun_out/r10_bench_strict.json | python tools/short.py r10_strict | tee -a gpurun_out/r10_run10.txt
