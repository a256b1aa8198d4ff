#!/bin/bash
S=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_gram.py -m gpu -q -s 2>&1 | grep -E "gram 358|passed|failed|Error|assert" | head
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "tests took $(( $(date +%s) - S )) s"
