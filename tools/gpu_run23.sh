#!/bin/bash
mkdir -p gpurun_out
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity --no-cpu-baseline"
for v in predA predB; do
CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so python bench.py $q 2>gpurun_out/r23_$v.err | python tools/short.py r23_$v | tee -a gpurun_out/r23_run23.txt
done
m="--workload ot --global-size 256 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline --single-mode"
for v in mhdD; do
CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so python bench.py $m 2>gpurun_out/r23_$v.err | python tools/short.py r23_$v | tee -a gpurun_out/r23_run23.txt
done
