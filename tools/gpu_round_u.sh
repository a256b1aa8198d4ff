#!/bin/bash
S=$(date +%s)
timeout 300 python tools/e2e_timing.py 2>&1 | tail -9
timeout 1200 python -m pytest tests/test_gpu_fit.py -m gpu -q 2>&1 | tail -3
echo "total $(( $(date +%s) - S )) s"
