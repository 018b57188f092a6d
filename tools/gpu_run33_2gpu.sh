#!/bin/bash
# 2-GPU call on the shipped state: rank-count independence, hydro with the overlapped exchange and MHD
mkdir -p gpurun_out
( time timeout 120 python -m pytest tests/test_gpu_multi.py -q -k "test_rank_count_independence[2] or test_rank_count_independence_mhd[2]" 2>&1 | tail -n 8 ) 2>&1 | tee gpurun_out/r5_pytest_multi2.log
