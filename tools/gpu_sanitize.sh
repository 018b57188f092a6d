#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_new_rows.py > gpurun_out/r4_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r4_memcheck.log
tail -6 gpurun_out/r4_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_new_rows.py > gpurun_out/r4_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r4_racecheck.log
tail -6 gpurun_out/r4_racecheck.log
