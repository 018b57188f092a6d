#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_decks.py tests/test_gpu_mhd.py tests/test_gpu_fullsize.py -q 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/r16_pytest.log
m="--workload ot --global-size 256 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline --single-mode"
python bench.py $m 2>gpurun_out/r16_ot.err | python tools/short.py ot256_ppmcall | tee -a gpurun_out/r16_run16.txt
