#!/bin/bash
mkdir -p gpurun_out
for v in "" f12 f16; do
  if [ -n "$v" ]; then export CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so; else unset CHOLLA_B200_LIB; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --precision 1 2>/dev/null | python scratch/short.py "f32_$v" | tee -a gpurun_out/f32b.txt
done
