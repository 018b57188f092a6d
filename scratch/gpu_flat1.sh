#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x > gpurun_out/pytest_flat1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_flat1.log
tail -3 gpurun_out/pytest_flat1.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
short() { python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('$1', 'Gcell/s %.3f'%(d['value']/1e9), 'ms/step %.2f'%d['ms_per_step'], 'pred %.2f corr %.2f'%(r['predictor_ms_per_launch'], r['ms_per_launch']))
"; }
for v in ""; do
  for m in fast strict; do CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200$v.so $B --mode $m 2>&1 | short L${v}_$m | tee -a gpurun_out/flat1.txt; done
done
