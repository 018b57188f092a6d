#!/bin/bash
mkdir -p gpurun_out
for w in kh noise; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --workload $w 2>gpurun_out/bench_wl_$w.err | tee gpurun_out/bench_wl_$w.json | python scratch/short.py "wl_$w" | tee -a gpurun_out/wl.txt
  tail -2 gpurun_out/bench_wl_$w.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_wl_ref_blast.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_wl_ref_blast.json').read()); print('ref blast', d['value']/1e9, d['ms_per_step'])"
