#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_decks.py -q -k "float32" 2>&1 | grep -v "^$" | tail -30
