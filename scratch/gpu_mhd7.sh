#!/bin/bash
mkdir -p gpurun_out
for v in nopark t64m8; do
  export CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so
  echo "variant [$v]"; python bench_mhd.py 256 5 | grep -v reference | tee -a gpurun_out/bench_mhd7.txt
done
