#!/bin/bash
# shipped state of the round: quick timing, the whole GPU test suite, the three bench arms as the driver runs them, profiles
mkdir -p gpurun_out /tmp/prof
C=$(cat tools/.commit 2>/dev/null || echo unknown)
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity --no-cpu-baseline"
python bench.py $q 2>gpurun_out/r32_quick.err | python tools/short.py r32_quick | tee -a gpurun_out/r32_run32.txt
m="--workload ot --global-size 256 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline --single-mode"
python bench.py $m 2>gpurun_out/r32_ot.err | python tools/short.py r32_ot256 | tee -a gpurun_out/r32_run32.txt
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 ) 2>&1 | tee gpurun_out/r5_pytest_all.log
( time python bench.py --impl reference ) > gpurun_out/r5_bench_ref.json 2> gpurun_out/r5_bench_ref.err
( time python bench.py ) > gpurun_out/r5_bench_full.json 2> gpurun_out/r5_bench_full.err
python tools/short.py full < gpurun_out/r5_bench_full.json | tee -a gpurun_out/r32_run32.txt
( time python bench.py --impl dropin ) > gpurun_out/r5_bench_dropin.json 2> gpurun_out/r5_bench_dropin.err
python - <<'PY'
import json
for f in ("ref", "dropin", "full"):
    try:
        d = json.loads([l for l in open("gpurun_out/r5_bench_%s.json" % f) if l.startswith("{")][-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("checksum"), (d.get("parity") or {}).get("bit_equal_vs_reference"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vl_dir_kernel -s 18 -c 6 -f -o /tmp/prof/r5_prof_strict_512 python bench.py --steps 1 --warmup 3 --mode strict --single-mode --no-cpu-baseline --no-e2e --no-configs --no-parity > gpurun_out/r5_prof_strict.log 2>&1
python tools/ncu_summary.py /tmp/prof/r5_prof_strict_512.ncu-rep > gpurun_out/r5_prof_strict_512.md
python tools/ncu_traffic.py 134217728 $C strict=/tmp/prof/r5_prof_strict_512.ncu-rep > gpurun_out/r5_roofline_traffic.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5_launches_bench_512.csv python bench.py --steps 2 --warmup 3 --single-mode --no-cpu-baseline --no-e2e --no-configs --no-parity > gpurun_out/r5_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mhd_ -s 33 -c 11 -f -o /tmp/prof/r5_prof_mhd_strict_256 python bench.py --workload ot --global-size 256 --steps 1 --warmup 3 --mode strict --single-mode --no-cpu-baseline --no-e2e --no-configs --no-parity > gpurun_out/r5_prof_mhd.log 2>&1
python tools/ncu_summary.py /tmp/prof/r5_prof_mhd_strict_256.ncu-rep > gpurun_out/r5_prof_mhd_strict_256.md
du -sh gpurun_out
