#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_split.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_split.log
tail -15 gpurun_out/pytest_split.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
short() { python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('$1', 'Gcell/s %.3f'%(d['value']/1e9), 'ms/step %.2f'%d['ms_per_step'], 'pred %.2f corr %.2f'%(r['predictor_ms_per_launch'], r['ms_per_launch']), d['clocks'])
"; }
$B --mode fast 2>&1 | short split_minb4_fast
$B --mode strict 2>&1 | short split_minb4_strict
for m in 5 6 8; do CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_minb$m.so $B --mode fast 2>&1 | short split_minb${m}_fast; done
CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_minb6.so $B --mode strict 2>&1 | short split_minb6_strict
CHL_SWEEP_IMPL=fused $B --mode fast 2>&1 | short fused_fast
for c in 16 32 128; do CHL_DIR_CHUNK=$c $B --mode fast 2>&1 | short split_minb4_fast_chunk$c; done
$B --mode fast --precision 1 2>&1 | short split_minb4_f32
