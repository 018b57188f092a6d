#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -q -k "exact_solver" 2>&1 | tail -25
