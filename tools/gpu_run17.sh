#!/bin/bash
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests/test_gpu_passive.py tests/test_gpu_cooling.py tests/test_gpu_dropin.py tests/test_gpu_mhd.py -q 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/r17_pytest.log
m="--workload ot --global-size 256 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline --single-mode"
python bench.py $m 2>gpurun_out/r17_ot.err | python tools/short.py ot256 | tee -a gpurun_out/r17_run17.txt
