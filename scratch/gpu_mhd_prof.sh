#!/bin/bash
mkdir -p gpurun_out
MHD_ARM=fast timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum --clock-control none -k regex:"mhd_|ghost" -s 60 -c 24 --csv --log-file gpurun_out/mhd_launches_fast.csv python bench_mhd.py 256 2 > gpurun_out/mhd_prof.log 2>&1
MHD_ARM=strict timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:"mhd_|ghost" -s 60 -c 24 --csv --log-file gpurun_out/mhd_launches_strict.csv python bench_mhd.py 256 2 >> gpurun_out/mhd_prof.log 2>&1
ls -la gpurun_out/mhd_launches*.csv
