#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parameters.py tests/test_gpu_parity.py -q -k "snapshot or deck or restart" 2>&1 | tail -15
