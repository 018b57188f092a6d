#!/bin/bash
# first GPU call of the round: bit-parity of the branch-free strict arithmetic, then its speed at several occupancies
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
timeout 1500 python -m pytest tests/test_gpu_exact_div.py tests/test_gpu_parity.py tests/test_gpu_reference.py -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_a.log
tail -3 gpurun_out/r2_pytest_a.log
run() { # tag, extra args
  tag=$1; shift
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --single-mode "$@" 2>gpurun_out/r2_bench_$tag.err | tee gpurun_out/r2_bench_$tag.json | python tools/short.py $tag | tee -a gpurun_out/r2_run1.txt
}
unset CHOLLA_B200_LIB
run default_strict --mode strict
run default_fast --mode fast
for v in s4 s5 s6 s8; do
  export CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so
  run $v --mode strict
done
unset CHOLLA_B200_LIB
run kh_strict --mode strict --workload kh
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_full.log
tail -3 gpurun_out/r2_pytest_full.log
