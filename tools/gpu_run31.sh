#!/bin/bash
# 32-bit cell index + per-use stencil offsets (variant idx) against the product library (one-test quotients + positive-pressure hint)
mkdir -p gpurun_out
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-cpu-baseline"
python bench.py $q --no-parity 2>gpurun_out/r31_default.err | python tools/short.py r31_default | tee -a gpurun_out/r31_run31.txt
for v in idx; do
CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so python bench.py $q 2>gpurun_out/r31_$v.err | tee gpurun_out/r31_$v.json | python tools/short.py r31_$v | tee -a gpurun_out/r31_run31.txt
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r31_$v.json") if l.startswith("{")][-1])
print("parity", d.get("parity", {}).get("bit_equal_vs_reference"))
PY
done
