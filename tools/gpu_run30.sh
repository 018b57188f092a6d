#!/bin/bash
# one range test per quotient (exact_div.cuh): timing of the headline and of Orszag-Tang, then the bit-parity suites
mkdir -p gpurun_out
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-cpu-baseline"
python bench.py $q 2>gpurun_out/r30_bench.err | tee gpurun_out/r30_bench.json | python tools/short.py r30_onetest | tee -a gpurun_out/r30_run30.txt
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r30_bench.json") if l.startswith("{")][-1])
print("parity", d.get("parity", {}).get("bit_equal_vs_reference"))
PY
m="--workload ot --global-size 256 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline --single-mode"
python bench.py $m 2>gpurun_out/r30_ot.err | python tools/short.py r30_ot256 | tee -a gpurun_out/r30_run30.txt
( time timeout 1500 python -m pytest tests/test_gpu_exact_div.py tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_mhd.py -q -x 2>&1 | tail -n 12 ) 2>&1 | tee gpurun_out/r30_pytest.log
