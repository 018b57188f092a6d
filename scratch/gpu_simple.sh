#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference.py tests/test_gpu_dropin.py -q -k "simple" 2>&1 | tail -25
