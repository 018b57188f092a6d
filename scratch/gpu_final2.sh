#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final2.log
tail -6 gpurun_out/pytest_final2.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; python scratch/short.py final2 < gpurun_out/bench_final2.json; tail -2 gpurun_out/bench_final2.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final2_ref.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_final2_ref.json').read()); print('ref', d['value']/1e9, d['ms_per_step'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --recon plmc --integrator simple 2>/dev/null | tee gpurun_out/bench_simple.json | python scratch/short.py simple_plmc_hllc | tee -a gpurun_out/combo.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 --recon plmc --integrator simple 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ref simple plmc hllc', d.get('value',0)/1e9, d.get('ms_per_step'), d.get('unavailable'))" | tee -a gpurun_out/combo.txt
