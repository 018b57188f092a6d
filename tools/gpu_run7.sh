#!/bin/bash
# 1-GPU call: fused predictor -- bits (vs the three sweeps, vs the oracle, vs the reference kernels), then speed
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -x -q 2>&1 | tail -8 | tee gpurun_out/r7_pytest.log
run() { tag=$1; shift
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --single-mode --no-configs --no-parity "$@" 2>gpurun_out/r7_bench_$tag.err | tee gpurun_out/r7_bench_$tag.json | python tools/short.py $tag | tee -a gpurun_out/r7_run7.txt; }
run fused_strict --mode strict
CHOLLA_SPLIT_PREDICTOR=1 run split_strict --mode strict
run fused_fast --mode fast
CHOLLA_SPLIT_PREDICTOR=1 run split_fast --mode fast
run fused_f32 --precision 1
run fused_kh --mode strict --workload kh
run fused_roe --mode strict --solver roe
run fused_256 --mode strict --size 256
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vl_predictor_kernel -s 3 -c 1 -f -o gpurun_out/r7_prof_predictor python bench.py --steps 1 --warmup 3 --mode strict --single-mode --no-cpu-baseline --no-e2e --no-configs --no-parity > gpurun_out/r7_prof.log 2>&1
ls -la gpurun_out/r7_prof_predictor.ncu-rep
