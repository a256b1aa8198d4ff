#!/bin/bash
# FP64 Cholesky with long rows cut into units (CMFB200_DMMA_SPLIT=1): parity, then config 3 with and without; e2e of the default workload
mkdir -p gpurun_out
CMFB200_DMMA_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_bench_shapes.py tests/test_gpu_sweeps.py -q -m gpu \
  -k "float64 or config3 or collective" 2>&1 | tail -5 > gpurun_out/r2s3_split_t1.log
cat gpurun_out/r2s3_split_t1.log
CMFB200_DMMA_SPLIT=1 timeout 300 python bench.py --workload ml10m_explicit_chol_k128_f64_sideinfo --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/r2s3_bench_cfg3_split.json 2> gpurun_out/r2s3_bench_cfg3_split.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2s3_bench_default2.json 2> gpurun_out/r2s3_bench_default2.err
timeout 300 python tools/e2e_timing.py 2>&1 | tail -13 > gpurun_out/r2s3_e2e_timing2.log
cat gpurun_out/r2s3_e2e_timing2.log
python - <<'PY'
import json
for f in ("r2s3_bench_cfg3_split", "r2s3_bench_default2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["roofline"].get("B_sweep_ms"), d["roofline"].get("A_sweep_ms"), d["e2e"])
    except Exception as e:
        print(f, "failed", e)
PY
