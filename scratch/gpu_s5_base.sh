#!/bin/bash
# session-5 baseline: full GPU suite, smoke, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_s5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s5.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_s5.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_s5.log
timeout 600 python bench.py > gpurun_out/bench_s5_default.json 2> gpurun_out/bench_s5_default.err
tail -5 gpurun_out/pytest_gpu_s5.log; tail -2 gpurun_out/smoke_s5.log; cat gpurun_out/bench_s5_default.json
