#!/bin/bash
mkdir -p gpurun_out/r3
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -60 > gpurun_out/r3/test_all.log; tail -25 gpurun_out/r3/test_all.log
echo "tests took $(( $(date +%s) - S )) s"
timeout 600 python tools/diag_implicit.py 2>&1 | tee gpurun_out/r3/diag_implicit.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg_resident_kernel -s 4 -c 2 -o gpurun_out/r3/full_resident_ml10m -f \
   python tools/quick_bench.py --shape ml10m --k 64 --iters 1 > gpurun_out/r3/ncu_full.log 2>&1
tail -3 gpurun_out/r3/ncu_full.log
echo "total $(( $(date +%s) - S )) s"
