#!/bin/bash
# register budget of the corrector passes chosen per direction (x / y / last pass); timing only
mkdir -p gpurun_out
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity --no-cpu-baseline"
for v in last6 l6y6 l6x6y6 l6x8 l6x6; do
CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so python bench.py $q 2>gpurun_out/r28_$v.err | python tools/short.py r28_$v | tee -a gpurun_out/r28_run28.txt
done
