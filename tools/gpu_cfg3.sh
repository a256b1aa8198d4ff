#!/bin/bash
# collective / side-information parity after the long-row fix of the FP64 Cholesky build kernel, then config 3 and config 4 lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fit.py tests/test_gpu_bench_shapes.py tests/test_gpu_foldin.py tests/test_gpu_sweeps.py -q -m gpu \
  -k "collective or side_information or precomputed or foldin or fold or float64 or chol" 2>&1 | tail -8 > gpurun_out/r2s3_cfg3_t1.log
cat gpurun_out/r2s3_cfg3_t1.log
timeout 300 python bench.py --workload ml10m_explicit_chol_k128_f64_sideinfo --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/r2s3_bench_cfg3.json 2> gpurun_out/r2s3_bench_cfg3.err
python - <<'PY'
import json
for f in ("r2s3_bench_cfg3",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["roofline"])
    except Exception as e:
        print(f, "failed", e)
PY
