#!/bin/bash
mkdir -p gpurun_out
# per step: ghost x6, predictor x3, corrector x3.  warm-up 3 steps (+ initial calc_dti + 6 ghost) then capture one step's six sweeps
for v in minb8 default; do
  if [ $v = default ]; then unset CHOLLA_B200_LIB; else export CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:vl_dir_kernel -s 18 -c 6 -o gpurun_out/prof_split_${v}_r1 python bench.py --steps 1 --warmup 3 --size 256 --mode fast --no-cpu-baseline --no-e2e > gpurun_out/prof_split_$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
