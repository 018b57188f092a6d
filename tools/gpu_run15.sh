#!/bin/bash
mkdir -p gpurun_out
( time timeout 2700 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/r15_pytest_all.log
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity --no-cpu-baseline"
python bench.py $q 2>gpurun_out/r15_bench_strict.err | tee gpurun_out/r15_bench_strict.json | python tools/short.py r15_strict | tee -a gpurun_out/r15_run15.txt
