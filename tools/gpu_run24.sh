#!/bin/bash
mkdir -p gpurun_out
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-cpu-baseline"
python bench.py $q 2>gpurun_out/r24_bench.err | tee gpurun_out/r24_bench.json | python tools/short.py r24_sidecache | tee -a gpurun_out/r24_run24.txt
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r24_bench.json") if l.startswith("{")][-1])
print("parity", d.get("parity", {}).get("bit_equal_vs_reference"), d["config"]["workload"])
PY
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r24_pytest.log
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('ref', d['value'], d['config'])"
