#!/bin/bash
# collective model after the tcgen05 Gram / lane-per-row triangular solve: parity, then the config-4 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_bench_shapes.py tests/test_gpu_foldin.py -q -m gpu -x \
  -k "collective or side_information or precomputed or foldin or fold" 2>&1 | tail -6 > gpurun_out/r2s3_cfg4_t1.log
cat gpurun_out/r2s3_cfg4_t1.log
timeout 300 python bench.py --workload ml10m_explicit_cg_k64_f32_implicit_features --no-cpu-baseline > gpurun_out/r2s3_bench_cfg4.json 2> gpurun_out/r2s3_bench_cfg4.err
tail -c 600 gpurun_out/r2s3_bench_cfg4.json
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2s3_bench_default.json 2> gpurun_out/r2s3_bench_default.err
python - <<'PY'
import json
for f in ("r2s3_bench_cfg4", "r2s3_bench_default"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"])
    except Exception as e:
        print(f, "failed", e)
PY
