#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --precision 1 2>gpurun_out/bench_f32.err | tee gpurun_out/bench_f32.json | python scratch/short.py f32
tail -3 gpurun_out/bench_f32.err
timeout 600 python -m pytest tests/test_gpu_mhd.py tests/test_gpu_exact_div.py -q -x 2>&1 | tail -3
