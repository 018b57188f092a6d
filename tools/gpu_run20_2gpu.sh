#!/bin/bash
# 2-GPU call: rank-count independence at 2 ranks for every transport / context, incl. the -DDE -DSCALAR -DSTATIC_GRAV one
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_multi.py -q -k "2" 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r20_pytest_multi2.log
