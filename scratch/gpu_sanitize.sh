#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scratch/sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize case done|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
