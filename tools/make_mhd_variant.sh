#!/bin/bash
# usage: tools/make_mhd_variant.sh NAME "<extra nvcc flags for the strict MHD TU>"
# Builds cholla_b200/libcholla_b200_NAME.so = the product library with its f64-strict MHD object recompiled with other
# flags.  Select with CHOLLA_B200_LIB.
set -e
cd "$(dirname "$0")/../cholla_b200/csrc"
name=$1; sflags=$2
B=build_var_$name; mkdir -p $B
COMMON="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr"
nvcc $COMMON -fmad=false -DCHL_STRICT=1 -DCHL_REAL=double -DCHL_MODE=f64_strict $sflags -Xptxas -v -c mhd_inst.cu -o $B/mhd_f64_strict.o 2> $B/ptxas_mhd_f64_strict.log
objs=$(ls build/*.o | grep -v mhd_f64_strict.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcholla_b200_$name.so $objs $B/mhd_f64_strict.o
grep -A2 "mhd_dir_kernel" $B/ptxas_mhd_f64_strict.log | grep -E "spill|Used" | paste - - | awk '{print $5, $9, $13, "regs", $(NF-12)}' | sort | uniq -c | head -12
