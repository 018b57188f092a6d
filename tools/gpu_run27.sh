#!/bin/bash
# prefetch of the predictor's next cell (pf: 7 CTAs, pf6: 6 CTAs / 80 registers) and a separate register budget for the
# last corrector pass (last6: 80 registers, last5: 96) against the shipped library; timing only
mkdir -p gpurun_out
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity --no-cpu-baseline"
python bench.py $q 2>gpurun_out/r27_default.err | python tools/short.py r27_default | tee -a gpurun_out/r27_run27.txt
for v in pf pf6 last6 last5; do
CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so python bench.py $q 2>gpurun_out/r27_$v.err | python tools/short.py r27_$v | tee -a gpurun_out/r27_run27.txt
done
