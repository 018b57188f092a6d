#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_decks.py -q -k "slice" 2>&1 | grep -v "^$" | tail -30
