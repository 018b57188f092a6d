#!/bin/bash
# validation pass after re-entry: parity tests, smoke, bench (strict, fast), reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v1.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_v1.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v1_strict.json 2> gpurun_out/bench_v1_strict.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode fast --no-cpu-baseline > gpurun_out/bench_v1_fast.json 2> gpurun_out/bench_v1_fast.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_v1_ref.json 2> gpurun_out/bench_v1_ref.err
tail -5 gpurun_out/pytest_gpu_v1.log; tail -2 gpurun_out/smoke_v1.log; cat gpurun_out/bench_v1_*.json
