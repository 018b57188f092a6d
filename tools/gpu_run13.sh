#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_cooling.py -q 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/r13_pytest_cooling.log
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity --no-cpu-baseline"
CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_ppmcall.so python bench.py $q 2>gpurun_out/r13_ppmcall.err | python tools/short.py r13_ppmcall | tee -a gpurun_out/r13_run13.txt
