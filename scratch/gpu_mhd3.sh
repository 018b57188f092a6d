#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mhd.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -3
python bench_mhd.py 256 5 | tee -a gpurun_out/bench_mhd3.txt
