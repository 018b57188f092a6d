#!/bin/bash
timeout 60 python -m pytest tests/test_gpu_decks.py tests/test_parameters.py -q -k "reference_ or restart or deck_driven" 2>&1 | grep -v "^$" | tail -25
