#!/bin/bash
# usage: tools/make_variant.sh NAME "<extra nvcc flags for the strict sweep TU>" [fast-flags]
# Builds cholla_b200/libcholla_b200_NAME.so = the product library with its f64-strict (and optionally f64-fast) sweep
# object recompiled with other flags (headline PPMC+HLLC subset only: -DCHL_DEV_SUBSET).  Select with CHOLLA_B200_LIB.
set -e
cd "$(dirname "$0")/../cholla_b200/csrc"
name=$1; sflags=$2; fflags=$3
B=build_var_$name; mkdir -p $B
COMMON="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -DCHL_DEV_SUBSET"
nvcc $COMMON -fmad=false -DCHL_STRICT=1 -DCHL_REAL=double -DCHL_MODE=f64_strict $sflags -Xptxas -v -c vl_sweep_inst.cu -o $B/sweep_f64_strict.o 2> $B/ptxas_f64_strict.log
objs="$(ls build/*.o | grep -v -e 'build/sweep_f64_strict.o' -e 'build/sweep_f64_fast.o' | tr '\n' ' ') $B/sweep_f64_strict.o"
if [ -n "$fflags" ]; then
  nvcc $COMMON -DCHL_STRICT=0 -DCHL_REAL=double -DCHL_MODE=f64_fast $fflags -Xptxas -v -c vl_sweep_inst.cu -o $B/sweep_f64_fast.o 2> $B/ptxas_f64_fast.log
  objs="$objs $B/sweep_f64_fast.o"
else
  objs="$objs build/sweep_f64_fast.o"
fi
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcholla_b200_$name.so $objs
grep -A2 "vl_dir_kernelILi2ELi0ELi1ELi2" $B/ptxas_f64_strict.log | grep -E "spill|Used" | tr '\n' ' '; echo
