#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference.py tests/test_gpu_parity.py -q -x -k "3d" 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --recon ppmp 2>/dev/null | python scratch/short.py ppmp_hllc | tee -a gpurun_out/combo.txt
python bench.py --impl reference --steps 3 --warmup 1 --recon ppmp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ref ppmp hllc', d.get('value',0)/1e9, d.get('ms_per_step'), d.get('unavailable'))" | tee -a gpurun_out/combo.txt
