#!/bin/bash
# bias initialisation: parity of the block-per-row path, then the ingestion stage timings with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py -q -m gpu -k "initial" -x 2>&1 | tail -5 > gpurun_out/r2s3_bias_t1.log
cat gpurun_out/r2s3_bias_t1.log
(echo "== CMFB200_BIAS_LONG=0"; CMFB200_BIAS_LONG=0 timeout 300 python tools/e2e_timing.py 2>&1 | tail -13; echo "== default"; timeout 300 python tools/e2e_timing.py 2>&1 | tail -13) > gpurun_out/r2s3_bias_timing.log 2>&1
cat gpurun_out/r2s3_bias_timing.log
