#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --solver hll 2>/dev/null | python scratch/short.py ppmc_hll | tee -a gpurun_out/combo.txt
python bench.py --impl reference --steps 3 --warmup 1 --solver hll 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ref ppmc hll', d.get('value',0)/1e9, d.get('ms_per_step'), d.get('unavailable'))" | tee -a gpurun_out/combo.txt
