#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact_div.py tests/test_gpu_reference.py tests/test_gpu_parity.py tests/test_gpu_mhd.py -q -x 2>&1 | tail -3
bash scratch/gpu_bench.sh q1
