#!/bin/bash
# strict PCM corrector on the conserved half-step state: bit parity against the reference's -DPCM builds (VL and SIMPLE)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_reference.py tests/test_gpu_parity.py tests/test_gpu_passive.py -q -x -k "pcm or PCM or strict_mode_is_bit_exact or passive_combinations or passive_3d or simple" 2>&1 | tail -n 15 ) 2>&1 | tee gpurun_out/r29_pytest.log
