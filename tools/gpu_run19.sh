#!/bin/bash
# final-state profiles, summarised ON the GPU box (the .ncu-rep files exceed what gpurun copies back): ncu --set full of one
# strict 512^3 step's six sweeps -> md summary + roofline_traffic.json, the launch list of the bench command, the MHD step
mkdir -p gpurun_out /tmp/prof
C=$(cat tools/.commit 2>/dev/null || echo unknown)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:vl_dir_kernel -s 18 -c 6 -f -o /tmp/prof/r4_prof_strict_512 python bench.py --steps 1 --warmup 3 --mode strict --single-mode --no-cpu-baseline --no-e2e --no-configs --no-parity > gpurun_out/r4_prof_strict.log 2>&1
python tools/ncu_summary.py /tmp/prof/r4_prof_strict_512.ncu-rep > gpurun_out/r4_prof_strict_512.md
python tools/ncu_traffic.py 134217728 $C strict=/tmp/prof/r4_prof_strict_512.ncu-rep > gpurun_out/r4_roofline_traffic.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4_launches_bench_512.csv python bench.py --steps 2 --warmup 3 --single-mode --no-cpu-baseline --no-e2e --no-configs --no-parity > gpurun_out/r4_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mhd_ -s 33 -c 11 -f -o /tmp/prof/r4_prof_mhd_strict_256 python bench.py --workload ot --global-size 256 --steps 1 --warmup 3 --mode strict --single-mode --no-cpu-baseline --no-e2e --no-configs --no-parity > gpurun_out/r4_prof_mhd.log 2>&1
python tools/ncu_summary.py /tmp/prof/r4_prof_mhd_strict_256.ncu-rep > gpurun_out/r4_prof_mhd_strict_256.md
ls -la /tmp/prof gpurun_out | tail -12
du -sh gpurun_out
