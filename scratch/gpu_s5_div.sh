#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact_div.py tests/test_gpu_reference.py tests/test_gpu_parity.py -q -x > gpurun_out/pytest_s5_div.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s5_div.log
tail -15 gpurun_out/pytest_s5_div.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_s5_div_strict.json 2> gpurun_out/bench_s5_div_strict.err
timeout 600 python bench.py --steps 5 --warmup 3 --mode fast --no-cpu-baseline --no-e2e > gpurun_out/bench_s5_div_fast.json 2> gpurun_out/bench_s5_div_fast.err
cat gpurun_out/bench_s5_div_strict.json gpurun_out/bench_s5_div_fast.json | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print(j['config']['workload'], 'ms/step', round(j['ms_per_step'],2), 'pred', round(j['roofline']['predictor_ms_per_launch'],2), 'corr', round(j['roofline']['ms_per_launch'],2), 'Gc/s', round(j['value']/1e9,3))
"
