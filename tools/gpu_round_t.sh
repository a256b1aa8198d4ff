#!/bin/bash
mkdir -p gpurun_out/r6
S=$(date +%s)
timeout 300 python tools/e2e_timing.py 2>&1 | tail -20
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r6/bench_default.json 2> gpurun_out/r6/bench_default.err; python -c "
import json; j=json.load(open('gpurun_out/r6/bench_default.json')); print(j['ms_per_step'], j['value'], j['e2e'], j['roofline']['frac'])"
echo "total $(( $(date +%s) - S )) s"
