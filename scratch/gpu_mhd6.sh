#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mhd.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -2
for v in "" mm5 mm6; do
  if [ -n "$v" ]; then export CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so; else unset CHOLLA_B200_LIB; fi
  echo "variant [$v]"; python bench_mhd.py 256 5 | grep -v reference | tee -a gpurun_out/bench_mhd6.txt
done
