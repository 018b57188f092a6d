#!/bin/bash
# usage: gpu_bench.sh tag [variant ...]   -- bench (both modes in one run) for the default library and variants
mkdir -p gpurun_out
tag=$1; shift
for v in "" "$@"; do
  if [ -n "$v" ]; then export CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so; else unset CHOLLA_B200_LIB; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_${tag}_$v.err | tee gpurun_out/bench_${tag}_$v.json | python scratch/short.py "${tag}_$v" | tee -a gpurun_out/$tag.txt
  tail -2 gpurun_out/bench_${tag}_$v.err
done
