#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r25_pytest.log
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-cpu-baseline"
python bench.py $q --workload kh --solver roe 2>gpurun_out/r25_kh_roe.err | tee gpurun_out/r25_kh_roe.json | python tools/short.py r25_kh_roe | tee -a gpurun_out/r25_run25.txt
python bench.py $q --workload kh 2>gpurun_out/r25_kh.err | tee gpurun_out/r25_kh.json | python tools/short.py r25_kh_hllc | tee -a gpurun_out/r25_run25.txt
python - <<'PY'
import json
for f in ("kh_roe", "kh"):
    d = json.loads([l for l in open("gpurun_out/r25_%s.json" % f) if l.startswith("{")][-1])
    print(f, "parity", d.get("parity", {}).get("bit_equal_vs_reference"))
PY
