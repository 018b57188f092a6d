#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_suite.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_suite.log
tail -5 gpurun_out/pytest_suite.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
