#!/bin/bash
S=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_gram.py -m gpu -q -x -s 2>&1 | tail -15
CMFB200_GRAM_TC=0 timeout 300 python -m pytest tests/test_gpu_gram.py -m gpu -q -x -s -k "matches" 2>&1 | tail -3
echo "tests took $(( $(date +%s) - S )) s"
