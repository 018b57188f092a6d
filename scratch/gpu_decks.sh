#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_decks.py -q 2>&1 | grep -v "^$" | grep "passed\|failed\|FAILED\|Error\|assert\|err" | head -40
