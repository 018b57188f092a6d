#!/bin/bash
mkdir -p gpurun_out
for c in "ppmc roe kh" "plmc hllc blast" "plmc roe blast"; do set -- $c
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --recon $1 --solver $2 --workload $3 2>gpurun_out/bench_combo_$1_$2.err | tee gpurun_out/bench_combo_$1_$2.json | python scratch/short.py "$1_$2_$3" | tee -a gpurun_out/combo.txt
  tail -1 gpurun_out/bench_combo_$1_$2.err
  python bench.py --impl reference --steps 3 --warmup 1 --recon $1 --solver $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ref $1 $2', d.get('value',0)/1e9, d.get('ms_per_step'), d.get('unavailable'))" | tee -a gpurun_out/combo.txt
done
