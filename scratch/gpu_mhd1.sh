#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_mhd.py -x -q > gpurun_out/pytest_mhd.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mhd.log
tail -40 gpurun_out/pytest_mhd.log
