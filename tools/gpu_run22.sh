#!/bin/bash
# the three bench arms as the driver runs them (defaults), one B200
mkdir -p gpurun_out
( time python bench.py --impl reference ) > gpurun_out/r4_bench_ref.json 2> gpurun_out/r4_bench_ref.err
( time python bench.py ) > gpurun_out/r4_bench_full.json 2> gpurun_out/r4_bench_full.err
tail -3 gpurun_out/r4_bench_full.err
python tools/short.py full < gpurun_out/r4_bench_full.json | tee -a gpurun_out/r4_run22.txt
( time python bench.py --impl dropin ) > gpurun_out/r4_bench_dropin.json 2> gpurun_out/r4_bench_dropin.err
python - <<'PY'
import json
for f in ("ref", "dropin"):
    d = json.loads([l for l in open("gpurun_out/r4_bench_%s.json" % f) if l.startswith("{")][-1])
    print(f, d.get("value"), d.get("ms_per_step"), d.get("checksum"))
PY
