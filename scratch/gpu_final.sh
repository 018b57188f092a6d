#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python scratch/short.py final < gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_final_ref.json').read()); print('ref', d['value']/1e9, d['ms_per_step'])"
timeout 600 python scratch/fast_vs_strict_blast.py 128 | tee gpurun_out/fast_vs_strict_blast.txt
python bench_mhd.py 256 5 | tee gpurun_out/bench_mhd_final.txt
