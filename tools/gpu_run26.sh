#!/bin/bash
mkdir -p gpurun_out
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-cpu-baseline"
python bench.py $q 2>gpurun_out/r26_bench.err | tee gpurun_out/r26_bench.json | python tools/short.py r26_epilogue | tee -a gpurun_out/r26_run26.txt
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r26_bench.json") if l.startswith("{")][-1])
print("parity", d.get("parity", {}).get("bit_equal_vs_reference"))
PY
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_decks.py tests/test_gpu_dropin.py tests/test_gpu_passive.py -q -x 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r26_pytest.log
