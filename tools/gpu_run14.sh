#!/bin/bash
mkdir -p gpurun_out
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity --no-cpu-baseline"
for v in evl ppmcall; do
CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so python bench.py $q 2>gpurun_out/r14_$v.err | python tools/short.py r14_$v | tee -a gpurun_out/r14_run14.txt
done
