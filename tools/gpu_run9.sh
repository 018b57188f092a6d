#!/bin/bash
# dual energy / scalars: parity tests first, then the whole GPU suite, then the headline line (no regression from the
# passive-field plumbing in the five-field kernels)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_passive.py -q 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/r9_pytest_passive.log
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity"
( time timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_passive.py 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r9_pytest_all.log
