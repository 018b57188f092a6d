#!/bin/bash
# first GPU pass: parity tests, smoke, bench (both modes), reference arm, ncu launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_strict.json 2> gpurun_out/bench_strict.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode fast --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode fast --precision 1 --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --impl reference --ref-variant nosync --steps 5 --warmup 1 > gpurun_out/bench_ref_nosync.json 2> gpurun_out/bench_ref_nosync.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --size 256 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench_*.json
